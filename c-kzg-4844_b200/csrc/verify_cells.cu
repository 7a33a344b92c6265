// verify_cell_kzg_proof_batch on the GPU.
//
// Replaces (paths relative to the reference tree), all of src/eip7594/eip7594.c:
//   compute_weighted_sum_of_commitments ................... :494-541
//   compute_commitment_to_aggregated_interpolation_poly ... :615-770
//   computed_weighted_sum_of_proofs ....................... :784-812
//   get_(inv_)coset_shift(_pow)_for_cell .................. :549-601
//   verify_cell_kzg_proof_batch (equation) ................ :825-974
//
//   e( sum_k r^k C_k  -  [I](tau)  +  sum_k r^k h_k^64 pi_k ,  G2 )  ==  e( sum_k r^k pi_k , [tau^64]G2 )
//
// Both sides are ONE pair of bucket MSMs over pre-shifted bases (vmsm.cu), as in the blob verifier:
//   * before the challenge r exists (and while the host still hashes the transcript): every proof and
//     every unique commitment is decompressed and subgroup-checked, and the doubling chains of that
//     check stay behind as the 18 table levels of the point (g1.cuh g1a_validate_levels);
//   * the 64 setup points the interpolation polynomial is committed with ([tau^j]G1, j < 64) have
//     their (negated) levels precomputed at setup (Ctx::mono_levels);
//   * after r: r^k, r^k h_k^64 (h_k = w8192^brp7(column k), eip7594.c:581-601), the commitment weights
//     w_c = sum_{k in c} r^k and the 64 interpolation coefficients are written as balanced base-|z|
//     digits; job A = sum r^k pi_k, job B = everything on the left -- no scalar multiplication and no
//     doubling after r.
// Grouping by column / by unique commitment is prepared on the host as CSR index lists (dedup order is
// part of the transcript, eip7594.c:345-376); the transcript itself is hashed on the host like the blob
// batch transcript (src/host_sha256.c).
#define KZG_FP_MUL_OUTLINE 1
#include <stdlib.h>
#include <string.h>

#include "cells.h"
#include "g1_glv.cuh"
#include "vmsm.cuh"

namespace kzg {

__device__ __forceinline__ uint32_t bswap32v(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ Fr ld_frv(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ int brp7v(int v) { return (int)(__brev((uint32_t)v) >> 25); }
__device__ __forceinline__ int brp6v(int v) { return (int)(__brev((uint32_t)v) >> 26); }

constexpr int VC_MONO = CELL_FR;  // 64 setup points carry the interpolation polynomial

// ---- before r: validation of the n proofs (columns 0..n-1) and u unique commitments (columns n..n+u-1)
__global__ void vc_validate_levels_kernel(G1* __restrict__ table, const uint8_t* __restrict__ proofs, uint64_t n, const uint8_t* __restrict__ uniq, uint64_t u, uint64_t npts,
                                          int* __restrict__ bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + u) return;
    const uint8_t* src = i < n ? proofs + 48 * i : uniq + 48 * (i - n);
    uint8_t buf[48];
    for (int q = 0; q < 48; q++) buf[q] = src[q];
    G1Affine a;
    if (!g1a_validate_levels(a, buf, table + i, npts)) *bad = 1;
}

// ---- after r ---------------------------------------------------------------------------------------
// rp[k] = r^k (Montgomery); digits of r^k (job A) and of r^k h_k^64 (job B, same column)
__global__ void vc_scalars_kernel(Fr* __restrict__ rp, uint32_t* __restrict__ hA, uint32_t* __restrict__ hB, const uint8_t* __restrict__ col_of, const Fr* __restrict__ roots,
                                  const Digest8 digest, uint32_t n) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Fr p = Fr::one();
    {
        Fr base = fr_from_digest_words(digest.h);
        uint32_t e = k;
        while (e) {
            if (e & 1) p = mul(p, base);
            base = sqr(base);
            e >>= 1;
        }
    }
    rp[k] = p;
    uint32_t s[8];
    from_mont<FrTag>(s, p);
    store_halves(hA, k, s);
    // h_k^64 = roots[64 * brp7(col)]  (eip7594.c:581-601)
    from_mont<FrTag>(s, mul(p, ld_frv(roots + 64 * brp7v(col_of[k]))));
    store_halves(hB, k, s);
}

// w_c = sum_{k in group c} rp[k]  (eip7594.c:523-531): one warp per unique commitment
__global__ void __launch_bounds__(32) vc_group_weights_kernel(uint32_t* __restrict__ hB, uint32_t first_point, const Fr* __restrict__ rp, const uint32_t* __restrict__ grp_start,
                                                              const uint32_t* __restrict__ grp_items) {
    __shared__ Fr sh[32];
    const uint32_t c = blockIdx.x, t = threadIdx.x;
    Fr s = Fr::zero();
    for (uint32_t i = grp_start[c] + t; i < grp_start[c + 1]; i += 32) s = add(s, ld_frv(rp + grp_items[i]));
    sh[t] = s;
    __syncwarp();
    if (t == 0) {
        for (int i = 1; i < 32; i++) s = add(s, sh[i]);
        uint32_t v[8];
        from_mont<FrTag>(v, s);
        store_halves(hB, (size_t)first_point + c, v);
    }
}

// part[slice][col][j] = sum_{k in slice of col} rp[k] * cell_k[j]   (canonical check of every cell
// element, eip7594.c:660-687).  grid (128 columns, slices)
__global__ void __launch_bounds__(64) vc_aggregate_columns_kernel(Fr* __restrict__ part, const uint8_t* __restrict__ cells, const Fr* __restrict__ rp,
                                                                  const uint32_t* __restrict__ col_start, const uint32_t* __restrict__ col_items, int* __restrict__ bad) {
    const int col = blockIdx.x, j = threadIdx.x;
    const uint32_t lo = col_start[col], len = col_start[col + 1] - lo;
    const uint32_t t0 = lo + (uint32_t)(((uint64_t)len * blockIdx.y) / gridDim.y), t1 = lo + (uint32_t)(((uint64_t)len * (blockIdx.y + 1)) / gridDim.y);
    Fr s = Fr::zero();
    for (uint32_t t = t0; t < t1; t++) {
        const uint32_t k = col_items[t];
        const uint4* p = reinterpret_cast<const uint4*>(cells + (size_t)k * CELL_BYTES + 32 * j);
        uint4 hi = __ldg(p), lo4 = __ldg(p + 1);
        uint32_t e[8] = {bswap32v(lo4.w), bswap32v(lo4.z), bswap32v(lo4.y), bswap32v(lo4.x), bswap32v(hi.w), bswap32v(hi.z), bswap32v(hi.y), bswap32v(hi.x)};
        if (limbs_geq<8>(e, FR_MOD)) *bad = 1;
        // the raw element is the Montgomery form of e/R: the factor R is restored by the R^2-scaled
        // constant of the interpolation kernel
        s = add(s, mul(Fr::from_limbs(e), ld_frv(rp + k)));
    }
    part[((size_t)blockIdx.y * CELLS_EXT + col) * CELL_FR + j] = s;
}

// F[col][k] = h_col^-k * (1/64) * sum_i agg[col][brp6(i)] * w64^(-i k)      (eip7594.c:713-741)
// h_col^-1 = roots[8192 - brp7(col)], w64^-1 = roots[8192 - 128].  Unused columns give zero rows.
__global__ void __launch_bounds__(64) vc_interpolate_kernel(Fr* __restrict__ F, const Fr* __restrict__ part, int slices, const Fr* __restrict__ roots) {
    __shared__ Fr v[CELL_FR];
    const int col = blockIdx.x, k = threadIdx.x;
    {
        const int src = brp6v(k);  // natural order: v[i] = f(h w64^i)
        Fr a = Fr::zero();
        for (int sidx = 0; sidx < slices; sidx++) a = add(a, ld_frv(part + ((size_t)sidx * CELLS_EXT + col) * CELL_FR + src));
        v[k] = a;
    }
    __syncthreads();
    Fr s = Fr::zero();
    for (int i = 0; i < CELL_FR; i++) {
        int e = (i * k) & 63;  // w64^(-ik) = roots[8192 - 128 * (ik mod 64)]
        Fr t = v[i];
        if (e != 0) t = mul(t, ld_frv(roots + (N_EXT - 128 * e)));
        s = add(s, t);
    }
    // FR_INV_64 is the Montgomery form R/64; one more factor R turns the raw sums (plain integers) into
    // Montgomery form: mont_mul(x, R^2/64) = x R / 64
    s = mul(s, to_mont<FrTag>(FR_INV_64));
    // h^-k by square-and-multiply on the 6-bit exponent
    Fr hinv = ld_frv(roots + (N_EXT - brp7v(col)));
    Fr p = Fr::one();
    for (int b = 5; b >= 0; b--) {
        p = sqr(p);
        if ((k >> b) & 1) p = mul(p, hinv);
    }
    F[col * CELL_FR + k] = mul(s, p);
}

// coeff[k] = sum_col F[col][k] -> digits for the setup column k (the table holds -[tau^k]G1)
__global__ void vc_sum_columns_kernel(uint32_t* __restrict__ hB, uint32_t first_point, const Fr* __restrict__ F) {
    int k = threadIdx.x;
    Fr s = Fr::zero();
    for (int col = 0; col < CELLS_EXT; col++) s = add(s, ld_frv(F + col * CELL_FR + k));
    uint32_t v[8];
    from_mont<FrTag>(v, s);
    store_halves(hB, (size_t)first_point + k, v);
}

// ------------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------------
static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }
static int vc_slices(uint64_t n) {
    uint64_t s = (n / CELLS_EXT + 31) / 32;  // ~32 cells per CTA when the columns are evenly filled
    return s < 1 ? 1 : (s > 32 ? 32 : (int)s);
}

size_t verify_cells_table_points(uint64_t n, uint64_t u) { return (size_t)VMSM_LEVELS * (n + u + VC_MONO); }

int setup_verify_cells(Launch& L, Ctx* c) {
    KZG_CUDA_TRY(cudaMalloc((void**)&c->mono_levels, (size_t)VMSM_LEVELS * VC_MONO * sizeof(G1)));
    return launch_vmsm_point_levels(L, c->mono_levels, c->g1_monomial, VC_MONO, true);
}

int launch_verify_cells_validate(Launch& L, G1* table, const uint8_t* proofs48, uint64_t n, const uint8_t* uniq48, uint64_t u, int* d_bad) {
    const uint64_t npts = n + u + VC_MONO;
    if (npts >= (1ull << 30) / VMSM_LEVELS) return RET_ERROR;
    // quads of lanes while the batch fits one wave (1024 points: 2.14 -> 1.33 ms; 8 k: 1.7 ms; 33 k: 5.6 ms against
    // 3.1 ms with one lane per point, r02l)
    static const bool quads_ok = !(getenv("CKZG_B200_VALIDATE") && strcmp(getenv("CKZG_B200_VALIDATE"), "lane") == 0);
    if (quads_ok && n + u <= 8448) {
        int rc = launch_g1_validate_levels_ab(L, nullptr, proofs48, n, 0, nullptr, uniq48, u, n, table, npts, d_bad);
        if (rc) return rc;
    } else {
        vc_validate_levels_kernel<<<(unsigned)((n + u + 31) / 32), 32, 0, L.stream>>>(table, proofs48, n, uniq48, u, npts, d_bad);
        KZG_CUDA_TRY(cudaGetLastError());
        L.count(1, "g1_validate");
    }
    // the fixed columns: 18 rows of 64 points
    KZG_CUDA_TRY(cudaMemcpy2DAsync(table + n + u, npts * sizeof(G1), L.ctx->mono_levels, VC_MONO * sizeof(G1), VC_MONO * sizeof(G1), VMSM_LEVELS, cudaMemcpyDeviceToDevice,
                                   L.stream));
    return RET_OK;
}

size_t verify_cells_scratch_bytes(uint64_t n, uint64_t u) {
    const uint64_t npts = n + u + VC_MONO;
    return a256(n * sizeof(Fr)) + a256(4 * n * 8) + a256(4 * npts * 8) + a256((size_t)vc_slices(n) * CELLS_EXT * CELL_FR * sizeof(Fr)) + a256(CELLS_EXT * CELL_FR * sizeof(Fr)) +
           vmsm_job_bytes(4 * n) + vmsm_job_bytes(4 * npts);
}

// out2[0] = A = sum r^k pi_k ; out2[1] = B = sum w_c C_c - [I] + sum r^k h_k^64 pi_k
int launch_verify_cells(Launch& L, G1* out2, const G1* table, const uint8_t* cells, const uint8_t* digest32, const uint8_t* col_of, const uint32_t* col_start,
                        const uint32_t* col_items, const uint32_t* cm_start, const uint32_t* cm_items, uint64_t n, uint64_t u, int* d_bad, void* scratch) {
    Ctx* c = L.ctx;
    const uint64_t npts = n + u + VC_MONO;
    const int slices = vc_slices(n);
    uint8_t* ws = (uint8_t*)scratch;
    Fr* rp = (Fr*)ws; ws += a256(n * sizeof(Fr));
    uint32_t* hA = (uint32_t*)ws; ws += a256(4 * n * 8);
    uint32_t* hB = (uint32_t*)ws; ws += a256(4 * npts * 8);
    Fr* part = (Fr*)ws; ws += a256((size_t)slices * CELLS_EXT * CELL_FR * sizeof(Fr));
    Fr* F = (Fr*)ws; ws += a256(CELLS_EXT * CELL_FR * sizeof(Fr));
    VmsmJobs jobs;
    ws = vmsm_job_carve(jobs.j[0], ws, hA, 4 * n);
    ws = vmsm_job_carve(jobs.j[1], ws, hB, 4 * npts);

    vc_scalars_kernel<<<(unsigned)((n + 63) / 64), 64, 0, L.stream>>>(rp, hA, hB, col_of, c->roots, digest8_from_bytes(digest32), (uint32_t)n);
    KZG_CUDA_TRY(cudaGetLastError());
    vc_group_weights_kernel<<<(unsigned)u, 32, 0, L.stream>>>(hB, (uint32_t)n, rp, cm_start, cm_items);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "vc_scalars");
    vc_aggregate_columns_kernel<<<dim3(CELLS_EXT, slices), 64, 0, L.stream>>>(part, cells, rp, col_start, col_items, d_bad);
    KZG_CUDA_TRY(cudaGetLastError());
    vc_interpolate_kernel<<<CELLS_EXT, 64, 0, L.stream>>>(F, part, slices, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    vc_sum_columns_kernel<<<1, 64, 0, L.stream>>>(hB, (uint32_t)(n + u), F);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(3, "vc_interpolation");
    return launch_vmsm_jobs(L, out2, jobs, table, (uint32_t)npts);
}

}  // namespace kzg
