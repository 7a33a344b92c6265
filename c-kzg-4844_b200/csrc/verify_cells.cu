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
// h_k = w8192^brp7(column k) is shared by every cell of a column, so the two proof sums are formed
// from ONE scalar multiplication per cell: S_col = sum_{k in col} r^k pi_k, then sum_col S_col and
// sum_col [h_col^64] S_col (128 extra multiplications instead of n).  Grouping by column / by unique
// commitment is prepared on the host as CSR index lists (dedup order is part of the transcript,
// eip7594.c:345-376); the transcript itself is hashed on the host like the blob batch transcript
// (src/host_sha256.c).
#define KZG_FP_MUL_OUTLINE 1
#include "cells.h"
#include "g1_glv.cuh"
#include "verify.h"

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

// rp[k] = r^k (Montgomery) and its plain limbs
__global__ void vc_powers_kernel(Fr* __restrict__ rp, uint32_t* __restrict__ rp_plain, const Fr* __restrict__ r, uint64_t n) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Fr p = Fr::one(), base = *r;
    uint64_t e = k;
    while (e) {
        if (e & 1) p = mul(p, base);
        base = sqr(base);
        e >>= 1;
    }
    rp[k] = p;
    from_mont<FrTag>(rp_plain + 8 * k, p);
}

// w[c] = sum_{k in group c} rp[k]   (plain limbs) -- commitment weights, eip7594.c:523-531
__global__ void vc_group_weights_kernel(uint32_t* __restrict__ w_plain, const Fr* __restrict__ rp, const uint32_t* __restrict__ grp_start, const uint32_t* __restrict__ grp_items, uint64_t groups) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= groups) return;
    Fr s = Fr::zero();
    for (uint32_t t = grp_start[c]; t < grp_start[c + 1]; t++) s = add(s, rp[grp_items[t]]);
    from_mont<FrTag>(w_plain + 8 * c, s);
}

// T[k] = [s_k] P_k for affine points and plain scalars; a second (points, scalars, out) triple may ride
// in the same launch (the kernel is one latency-bound multiplication per thread)
__global__ void __launch_bounds__(64) vc_scalar_mul_kernel(G1* __restrict__ T, const G1Affine* __restrict__ P, const uint32_t* __restrict__ s_plain, uint64_t n,
                                                           G1* __restrict__ T2, const G1Affine* __restrict__ P2, const uint32_t* __restrict__ s2_plain, uint64_t n2) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n + n2) return;
    if (k >= n) {
        k -= n;
        T = T2;
        P = P2;
        s_plain = s2_plain;
    }
    uint32_t kk[8];
    for (int q = 0; q < 8; q++) kk[q] = s_plain[8 * k + q];
    T[k] = g1_mul_glv_affine(P[k], kk);
}

// S[col] = sum_{k in col} T[k];  H[col] = [h_col^64] S[col], h_col^64 = roots[64 * brp7(col)]  (eip7594.c:581-601)
__global__ void __launch_bounds__(64) vc_column_sums_kernel(G1* __restrict__ S, G1* __restrict__ H, const G1* __restrict__ T, const uint32_t* __restrict__ col_start,
                                                            const uint32_t* __restrict__ col_items, const Fr* __restrict__ roots) {
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= CELLS_EXT) return;
    G1 acc = g1_inf();
    for (uint32_t t = col_start[col]; t < col_start[col + 1]; t++) {
        G1 p = T[col_items[t]];
        g1_add_to(acc, p);
    }
    S[col] = acc;
    uint32_t k[8];
    from_mont<FrTag>(k, ld_frv(roots + 64 * brp7v(col)));
    G1 m = g1_mul_glv(acc, k);
    H[col] = m;
}

// agg[col][j] = sum_{k in col} rp[k] * cell_k[j]   (canonical check of every cell element, eip7594.c:660-687)
__global__ void __launch_bounds__(64) vc_aggregate_columns_kernel(Fr* __restrict__ agg, const uint8_t* __restrict__ cells, const Fr* __restrict__ rp, const uint32_t* __restrict__ col_start,
                                                                  const uint32_t* __restrict__ col_items, int* __restrict__ bad) {
    const int col = blockIdx.x, j = threadIdx.x;  // 128 x 64
    Fr s = Fr::zero();
    for (uint32_t t = col_start[col]; t < col_start[col + 1]; t++) {
        const uint32_t k = col_items[t];
        const uint4* p = reinterpret_cast<const uint4*>(cells + (size_t)k * CELL_BYTES + 32 * j);
        uint4 hi = __ldg(p), lo = __ldg(p + 1);
        uint32_t e[8] = {bswap32v(lo.w), bswap32v(lo.z), bswap32v(lo.y), bswap32v(lo.x), bswap32v(hi.w), bswap32v(hi.z), bswap32v(hi.y), bswap32v(hi.x)};
        if (limbs_geq<8>(e, FR_MOD)) *bad = 1;
        s = add(s, mul(to_mont<FrTag>(e), ld_frv(rp + k)));
    }
    agg[col * CELL_FR + j] = s;
}

// F[col][k] = h_col^-k * (1/64) * sum_i agg[col][brp6(i)] * w64^(-i k)      (eip7594.c:713-741)
// h_col^-1 = roots[8192 - brp7(col)], w64^-1 = roots[8192 - 128].  Unused columns give zero rows.
__global__ void __launch_bounds__(64) vc_interpolate_kernel(Fr* __restrict__ F, const Fr* __restrict__ agg, const Fr* __restrict__ roots) {
    __shared__ Fr v[CELL_FR];
    const int col = blockIdx.x, k = threadIdx.x;
    v[k] = agg[col * CELL_FR + brp6v(k)];  // natural order: v[i] = f(h w64^i)
    __syncthreads();
    Fr s = Fr::zero();
    for (int i = 0; i < CELL_FR; i++) {
        int e = (i * k) & 63;  // w64^(-ik) = roots[8192 - 128 * (ik mod 64)]
        Fr t = v[i];
        if (e != 0) t = mul(t, ld_frv(roots + (N_EXT - 128 * e)));
        s = add(s, t);
    }
    s = mul(s, Fr::from_limbs(FR_INV_64));
    // h^-k by square-and-multiply on the 6-bit exponent
    Fr hinv = ld_frv(roots + (N_EXT - brp7v(col)));
    Fr p = Fr::one();
    for (int b = 5; b >= 0; b--) {
        p = sqr(p);
        if ((k >> b) & 1) p = mul(p, hinv);
    }
    F[col * CELL_FR + k] = mul(s, p);
}

// coeff[k] = sum_col F[col][k]  (plain limbs)
__global__ void vc_sum_columns_kernel(uint32_t* __restrict__ coeff_plain, const Fr* __restrict__ F) {
    int k = threadIdx.x;
    Fr s = Fr::zero();
    for (int col = 0; col < CELLS_EXT; col++) s = add(s, F[col * CELL_FR + k]);
    from_mont<FrTag>(coeff_plain + 8 * k, s);
}

__global__ void vc_negate_kernel(G1* p) {
    if (threadIdx.x == 0 && blockIdx.x == 0) p->y = neg(p->y);
}

// ------------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------------
static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t verify_cells_scratch_bytes(uint64_t n, uint64_t u) {
    uint64_t fold = (n + u + 512) / 1024 + 4;
    return a256(n * sizeof(Fr)) + a256(n * 32) + a256(u * 32) + a256((n + fold) * sizeof(G1)) + a256((u + fold) * sizeof(G1)) + 3 * a256((CELLS_EXT + 4) * sizeof(G1)) +
           2 * a256(CELLS_EXT * CELL_FR * sizeof(Fr)) + a256(64 * 32) + a256((64 + 4) * sizeof(G1)) + a256(8 * sizeof(G1));
}

// out2[0] = A = sum r^k pi_k ; out2[1] = B = sum w_c C_c - [I] + sum r^k h_k^64 pi_k
int launch_verify_cells(Launch& L, G1* out2, const G1Affine* proofs, const G1Affine* commitments, const uint8_t* cells, const Fr* r, const uint32_t* col_start,
                        const uint32_t* col_items, const uint32_t* cm_start, const uint32_t* cm_items, uint64_t n, uint64_t u, int* d_bad, void* scratch) {
    Ctx* c = L.ctx;
    uint64_t fold = (n + u + 512) / 1024 + 4;
    uint8_t* ws = (uint8_t*)scratch;
    Fr* rp = (Fr*)ws; ws += a256(n * sizeof(Fr));
    uint32_t* rp_plain = (uint32_t*)ws; ws += a256(n * 32);
    uint32_t* w_plain = (uint32_t*)ws; ws += a256(u * 32);
    G1* T = (G1*)ws; ws += a256((n + fold) * sizeof(G1));
    G1* TC = (G1*)ws; ws += a256((u + fold) * sizeof(G1));
    G1* S = (G1*)ws; ws += a256((CELLS_EXT + 4) * sizeof(G1));
    G1* H = (G1*)ws; ws += a256((CELLS_EXT + 4) * sizeof(G1));
    G1* tmp = (G1*)ws; ws += a256((CELLS_EXT + 4) * sizeof(G1));
    Fr* agg = (Fr*)ws; ws += a256(CELLS_EXT * CELL_FR * sizeof(Fr));
    Fr* F = (Fr*)ws; ws += a256(CELLS_EXT * CELL_FR * sizeof(Fr));
    uint32_t* coeff = (uint32_t*)ws; ws += a256(64 * 32);
    G1* TI = (G1*)ws; ws += a256((64 + 4) * sizeof(G1));
    G1* parts = (G1*)ws;  // [0] sum_c, [1] -I, [2] sum H, then fold space
    (void)tmp;
    const unsigned nb64 = (unsigned)((n + 63) / 64), ub64 = (unsigned)((u + 63) / 64);

    vc_powers_kernel<<<nb64, 64, 0, L.stream>>>(rp, rp_plain, r, n);
    KZG_CUDA_TRY(cudaGetLastError());
    vc_group_weights_kernel<<<ub64, 64, 0, L.stream>>>(w_plain, rp, cm_start, cm_items, u);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "vc_scalars");
    // one scalar multiplication per cell proof and per unique commitment
    vc_scalar_mul_kernel<<<(unsigned)((n + u + 63) / 64), 64, 0, L.stream>>>(T, proofs, rp_plain, n, TC, commitments, w_plain, u);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "vc_scalar_mul");
    // The interpolation chain (aggregate -> 64-point inverse transforms -> 64 multiplications on the
    // monomial setup) does not depend on the proofs: it runs on a side stream unless per-kernel profiling
    // wants everything on one stream.
    cudaStream_t side = L.stream;
    const bool forked = (L.trace == nullptr);
    if (forked) {
        cudaEvent_t ev;
        if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) return RET_ERROR;
        KZG_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        KZG_CUDA_TRY(cudaEventRecord(ev, L.stream));  // rp[] is ready at this point
        cudaStreamWaitEvent(side, ev, 0);
        cudaEventDestroy(ev);
    }
    vc_aggregate_columns_kernel<<<CELLS_EXT, 64, 0, side>>>(agg, cells, rp, col_start, col_items, d_bad);
    KZG_CUDA_TRY(cudaGetLastError());
    vc_interpolate_kernel<<<CELLS_EXT, 64, 0, side>>>(F, agg, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    vc_sum_columns_kernel<<<1, 64, 0, side>>>(coeff, F);
    KZG_CUDA_TRY(cudaGetLastError());
    vc_scalar_mul_kernel<<<1, 64, 0, side>>>(TI, c->g1_monomial, coeff, 64, nullptr, nullptr, nullptr, 0);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(4, "vc_interpolation");
    vc_column_sums_kernel<<<CELLS_EXT / 64, 64, 0, L.stream>>>(S, H, T, col_start, col_items, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "vc_column_sums");
    if (forked) {
        cudaEvent_t done;
        KZG_CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        KZG_CUDA_TRY(cudaEventRecord(done, side));
        cudaStreamWaitEvent(L.stream, done, 0);
        cudaEventDestroy(done);
        cudaStreamDestroy(side);
    }
    int rc;
    if ((rc = launch_g1_sum(L, out2 + 0, S, CELLS_EXT))) return rc;  // A (S is clobbered: H was computed first)
    if ((rc = launch_g1_sum(L, parts + 0, TC, u))) return rc;
    if ((rc = launch_g1_sum(L, parts + 1, TI, 64))) return rc;
    vc_negate_kernel<<<1, 32, 0, L.stream>>>(parts + 1);
    KZG_CUDA_TRY(cudaGetLastError());
    if ((rc = launch_g1_sum(L, parts + 2, H, CELLS_EXT))) return rc;
    if ((rc = launch_g1_sum(L, out2 + 1, parts, 3))) return rc;
    return RET_OK;
}

}  // namespace kzg
