// Cell recovery (erasure decoding) for recover_cells_and_kzg_proofs.
//
// Replaces (paths relative to the reference tree):
//   recover_cells ............................... src/eip7594/recovery.c:200-365
//   vanishing_polynomial_for_missing_cells ...... src/eip7594/recovery.c:93-162 (+ :46-91)
//   coset_fft / coset_ifft / shift_poly ......... src/eip7594/fft.c:257-301, src/eip7594/poly.c:38
//
// Same mathematics as the reference, so the result is identical even for inconsistent inputs:
//   E*Z on the domain -> coefficients Q -> Q on the coset 7*<w> -> divide by Z on the coset ->
//   coefficients P -> evaluations.  B200 arrangement:
//   * Z(x) = Zs(x^64) with deg Zs = #missing cells <= 64, so Z takes ONE value per cell on the domain
//     and on the coset: 2 x 128 Horner evaluations replace the reference's two 8192-point transforms,
//     and the 8192 field inversions of recovery.c:322-328 become 128;
//   * cells arrive in bit-reversed evaluation order = the input order of a decimation-in-time inverse
//     transform, forward transforms run decimation-in-frequency and leave bit-reversed output: the four
//     8192-point transforms need no permutation pass at all;
//   * an 8192-point transform = one radix-2 stage in global memory (fused with the scalings, the coset
//     shift and the next transform's first stage) + two 4096-point transforms in shared memory.
#include "cells.h"

namespace kzg {

__device__ __forceinline__ uint32_t bswap32r(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ Fr ld_frr(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ int brp7r(int v) { return (int)(__brev((uint32_t)v) >> 25); }

// ------------------------------------------------------------------------------------------------
// setup-time tables: shiftA[k] = 7^k / 8192, shiftB[k] = 7^-k / 8192   (k < 8192)
// ------------------------------------------------------------------------------------------------
__global__ void recover_tables_kernel(Fr* __restrict__ shiftA, Fr* __restrict__ shiftB) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N_EXT) return;
    Fr s = Fr::from_limbs(FR_SHIFT), si = Fr::from_limbs(FR_SHIFT_INV);
    Fr a = Fr::one(), b = Fr::one();
    for (int bit = 12; bit >= 0; bit--) {
        a = sqr(a);
        b = sqr(b);
        if ((k >> bit) & 1) {
            a = mul(a, s);
            b = mul(b, si);
        }
    }
    Fr inv = Fr::from_limbs(FR_INV_8192);
    shiftA[k] = mul(a, inv);
    shiftB[k] = mul(b, inv);
}

// ------------------------------------------------------------------------------------------------
// per blob: short vanishing polynomial, its value per cell on the domain (zc) and the inverse of its
// value per cell on the coset (zq_inv)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) recover_vanishing_kernel(Fr* __restrict__ zc, Fr* __restrict__ zq_inv, const uint8_t* __restrict__ present, const Fr* __restrict__ roots) {
    __shared__ Fr poly[66];
    __shared__ Fr nxt[66];
    __shared__ int miss[128];
    __shared__ int n_miss;
    const int blob = blockIdx.x, t = threadIdx.x;
    const uint8_t* pr = present + (size_t)blob * 128;
    if (t == 0) {
        int m = 0;
        for (int c = 0; c < 128; c++)
            if (!pr[c]) miss[m++] = c;
        n_miss = m;
    }
    if (t < 66) poly[t] = (t == 0) ? Fr::one() : Fr::zero();
    __syncthreads();
    const int m = n_miss;  // <= 64 (checked on the host, eip7594.c:191-200)
    // poly <- poly * (x - r_i), r_i = w128^brp7(c_i) = roots[64 * brp7(c_i)]   (recovery.c:46-91,118-122)
    for (int i = 0; i < m; i++) {
        Fr r = ld_frr(roots + 64 * brp7r(miss[i]));
        if (t <= i + 1) {
            Fr lo = (t <= i) ? mul(poly[t], neg(r)) : Fr::zero();
            Fr hi = (t >= 1) ? poly[t - 1] : Fr::zero();
            nxt[t] = add(lo, hi);
        }
        __syncthreads();
        if (t <= i + 1) poly[t] = nxt[t];
        __syncthreads();
    }
    // Horner at w128^brp7(c) (domain) and at 7^64 * w128^brp7(c) (coset)
    {
        const int c = t;
        Fr x = ld_frr(roots + 64 * brp7r(c));
        Fr s64 = Fr::from_limbs(FR_SHIFT);
#pragma unroll 1
        for (int k = 0; k < 6; k++) s64 = sqr(s64);  // 7^64
        Fr xq = mul(x, s64);
        Fr a = poly[m], b = poly[m];
        for (int k = m - 1; k >= 0; k--) {
            a = add(mul(a, x), poly[k]);
            b = add(mul(b, xq), poly[k]);
        }
        zc[(size_t)blob * 128 + c] = a;
        zq_inv[(size_t)blob * 128 + c] = fr_inv(b);  // Z has no root on the coset: b != 0
    }
}

// A[blob][q] = E[q] * zc[cell(q)] for received cells (canonical check), 0 for missing ones.
// `slot[blob][c]` = position of cell c in the caller's list, or -1.
__global__ void recover_scatter_kernel(Fr* __restrict__ A, const uint8_t* __restrict__ cells, const int16_t* __restrict__ slot, const Fr* __restrict__ zc, uint64_t num_cells,
                                       int* __restrict__ bad) {
    const int blob = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;  // < 8192
    const int c = q >> 6, j = q & 63;
    const int sl = slot[(size_t)blob * 128 + c];
    Fr v = Fr::zero();
    if (sl >= 0) {
        const uint4* p = reinterpret_cast<const uint4*>(cells + ((size_t)blob * num_cells + sl) * CELL_BYTES + 32 * j);
        uint4 hi = __ldg(p), lo = __ldg(p + 1);
        uint32_t s[8] = {bswap32r(lo.w), bswap32r(lo.z), bswap32r(lo.y), bswap32r(lo.x), bswap32r(hi.w), bswap32r(hi.z), bswap32r(hi.y), bswap32r(hi.x)};
        if (limbs_geq<8>(s, FR_MOD)) bad[blob] = 1;  // bytes_to_bls_field (eip7594.c:233)
        v = mul(to_mont<FrTag>(s), ld_frr(zc + (size_t)blob * 128 + c));
    }
    A[(size_t)blob * N_EXT + q] = v;
}

// ------------------------------------------------------------------------------------------------
// 4096-point halves in shared memory
// ------------------------------------------------------------------------------------------------
constexpr int RH_THREADS = 512;

// INVERSE: decimation in time over A[blob][h*4096 ..], bit-reversed in -> natural out, optional
// per-cell multiplier applied while loading.  FORWARD: decimation in frequency, natural in ->
// bit-reversed out, optionally written as canonical big-endian bytes (the recovered cells).
template <bool INVERSE>
__global__ void __launch_bounds__(RH_THREADS) recover_half_kernel(Fr* __restrict__ A, const Fr* __restrict__ cell_mult, uint8_t* __restrict__ bytes_out, const Fr* __restrict__ roots) {
    extern __shared__ uint4 smem_raw[];
    Fr* sh = reinterpret_cast<Fr*>(smem_raw);
    const int blob = blockIdx.y, h = blockIdx.x, tid = threadIdx.x;
    Fr* base = A + (size_t)blob * N_EXT + (size_t)h * N_BLOB;
    for (int i = tid; i < N_BLOB; i += RH_THREADS) {
        Fr v = base[i];
        if (INVERSE && cell_mult) v = mul(v, ld_frr(cell_mult + (size_t)blob * 128 + ((h * N_BLOB + i) >> 6)));
        sh[i] = v;
    }
    __syncthreads();
    if (INVERSE) {
#pragma unroll 1
        for (int half = 1; half < N_BLOB; half <<= 1) {
            const int tw_step = (N_EXT / 2) / half;
#pragma unroll 1
            for (int b = tid; b < N_BLOB / 2; b += RH_THREADS) {
                int j = b & (half - 1);
                int i0 = ((b - j) << 1) + j, i1 = i0 + half;
                Fr u = sh[i0], v = sh[i1];
                if (j != 0) v = mul(v, ld_frr(roots + (N_EXT - j * tw_step)));
                sh[i0] = add(u, v);
                sh[i1] = sub(u, v);
            }
            __syncthreads();
        }
    } else {
#pragma unroll 1
        for (int half = N_BLOB / 2; half >= 1; half >>= 1) {
            const int tw_step = (N_EXT / 2) / half;
#pragma unroll 1
            for (int b = tid; b < N_BLOB / 2; b += RH_THREADS) {
                int j = b & (half - 1);
                int i0 = ((b - j) << 1) + j, i1 = i0 + half;
                Fr u = sh[i0], v = sh[i1];
                Fr d = sub(u, v);
                if (j != 0) d = mul(d, ld_frr(roots + j * tw_step));
                sh[i0] = add(u, v);
                sh[i1] = d;
            }
            __syncthreads();
        }
    }
    if (!INVERSE && bytes_out) {
        uint8_t* dst = bytes_out + (size_t)blob * 2 * BLOB_BYTES + (size_t)h * BLOB_BYTES;
        for (int i = tid; i < N_BLOB; i += RH_THREADS) {
            uint32_t t[8];
            from_mont<FrTag>(t, sh[i]);
            uint4* d = reinterpret_cast<uint4*>(dst + 32 * i);
            d[0] = make_uint4(bswap32r(t[7]), bswap32r(t[6]), bswap32r(t[5]), bswap32r(t[4]));
            d[1] = make_uint4(bswap32r(t[3]), bswap32r(t[2]), bswap32r(t[1]), bswap32r(t[0]));
        }
    } else {
        for (int i = tid; i < N_BLOB; i += RH_THREADS) base[i] = sh[i];
    }
}

// Outer stage of the inverse transform (natural output), coefficient scaling `shift[k]` (1/8192 and
// the coset (un)shift), optional copy of the low 4096 coefficients, then the outer stage of the next
// forward transform -- all on the pair (j, j + 4096) a thread already holds.
__global__ void recover_outer_kernel(Fr* __restrict__ A, const Fr* __restrict__ shift, Fr* __restrict__ mono_out, const Fr* __restrict__ roots) {
    const int blob = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;  // < 4096
    Fr* a = A + (size_t)blob * N_EXT;
    // inverse DIT last stage: twiddle w8192^-j
    Fr u = a[j], v = a[j + N_BLOB];
    if (j != 0) v = mul(v, ld_frr(roots + (N_EXT - j)));
    Fr x0 = mul(add(u, v), ld_frr(shift + j));
    Fr x1 = mul(sub(u, v), ld_frr(shift + j + N_BLOB));
    if (mono_out) mono_out[(size_t)blob * N_BLOB + j] = x0;
    // forward DIF first stage: twiddle w8192^j
    Fr d = sub(x0, x1);
    if (j != 0) d = mul(d, ld_frr(roots + j));
    a[j] = add(x0, x1);
    a[j + N_BLOB] = d;
}

int recover_setup(Launch& L, Ctx* c) {
    KZG_CUDA_TRY(cudaMalloc((void**)&c->rec_shiftA, N_EXT * sizeof(Fr)));
    KZG_CUDA_TRY(cudaMalloc((void**)&c->rec_shiftB, N_EXT * sizeof(Fr)));
    recover_tables_kernel<<<N_EXT / 128, 128, 0, L.stream>>>((Fr*)c->rec_shiftA, (Fr*)c->rec_shiftB);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaFuncSetAttribute(recover_half_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, N_BLOB * (int)sizeof(Fr)));
    KZG_CUDA_TRY(cudaFuncSetAttribute(recover_half_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, N_BLOB * (int)sizeof(Fr)));
    L.count(1, "recover_setup");
    return RET_OK;
}

// cells_out: n x 128 x 2048 bytes (device); mono_out: n x 4096 Fr (device, may be null);
// cells_in: n x num_cells x 2048 bytes; slot: n x 128 (position of each cell in the input or -1);
// present: n x 128 flags.  scratch: recover_scratch_bytes(n).
size_t recover_scratch_bytes(uint64_t n) { return n * N_EXT * sizeof(Fr) + 2 * n * 128 * sizeof(Fr) + 512; }

int launch_recover(Launch& L, uint8_t* cells_out, Fr* mono_out, const uint8_t* cells_in, const int16_t* slot, const uint8_t* present, uint64_t num_cells, uint64_t n, int* d_bad,
                   void* scratch) {
    if (!n) return RET_OK;
    Ctx* c = L.ctx;
    uint8_t* ws = (uint8_t*)scratch;
    Fr* A = (Fr*)ws;
    ws += n * N_EXT * sizeof(Fr);
    Fr* zc = (Fr*)ws;
    ws += n * 128 * sizeof(Fr);
    Fr* zq_inv = (Fr*)ws;
    const size_t smem = N_BLOB * sizeof(Fr);
    dim3 g_el(N_EXT / 256, (unsigned)n), g_half(2, (unsigned)n), g_outer(N_BLOB / 256, (unsigned)n);

    recover_vanishing_kernel<<<(unsigned)n, 128, 0, L.stream>>>(zc, zq_inv, present, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    recover_scatter_kernel<<<g_el, 256, 0, L.stream>>>(A, cells_in, slot, zc, num_cells, d_bad);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "recover_prepare");
    // (E*Z) -> Q (coefficients) -> Q on the coset
    recover_half_kernel<true><<<g_half, RH_THREADS, smem, L.stream>>>(A, nullptr, nullptr, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    recover_outer_kernel<<<g_outer, 256, 0, L.stream>>>(A, (const Fr*)c->rec_shiftA, nullptr, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    recover_half_kernel<false><<<g_half, RH_THREADS, smem, L.stream>>>(A, nullptr, nullptr, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    // / Z on the coset -> P (coefficients) -> P on the domain
    recover_half_kernel<true><<<g_half, RH_THREADS, smem, L.stream>>>(A, zq_inv, nullptr, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    recover_outer_kernel<<<g_outer, 256, 0, L.stream>>>(A, (const Fr*)c->rec_shiftB, mono_out, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    recover_half_kernel<false><<<g_half, RH_THREADS, smem, L.stream>>>(A, nullptr, cells_out, c->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(6, "recover_transforms");
    return RET_OK;
}

}  // namespace kzg
