// EIP-7594 scalar side: blob -> monomial coefficients -> extended evaluations (cells), and the
// 64 x FFT-128 of the FK20 circulant vectors.
//
// Replaces (paths relative to the reference tree):
//   poly_lagrange_to_monomial ............ src/eip7594/poly.c:58 (bit-reversal + fr_ifft 4096)
//   fr_fft / fr_ifft / fr_fft_fast ....... src/eip7594/fft.c:70-146
//   compute_cells_and_kzg_proofs (cells) . src/eip7594/eip7594.c:88-121
//   compute_fk20_cell_proofs, phase 1 .... src/eip7594/fk20.c:55-78,199-209
//
// One CTA per blob with the whole 4096-point polynomial (128 KiB) in shared memory, radix-2 stages.
// No permutation passes are needed:
//   * the blob is given in bit-reversed evaluation order, which is exactly the input order of a
//     decimation-in-time inverse transform producing natural-order coefficients a[m];
//   * the 8192 extended evaluations in bit-reversed order are [blob itself | NTT_4096(a[m] w8192^m)]
//     and a decimation-in-frequency transform of natural input leaves its output bit-reversed.
// So cells 0..63 are the (validated) input bytes and cells 64..127 one 4096-point DIF transform.
#include "cells.h"
#include "tma.cuh"

namespace kzg {

__device__ __forceinline__ uint32_t bswap32c(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

constexpr int NTT_THREADS = 512;
constexpr int NTT_N = 4096;

// In-place radix-2 stages over sh[0..4096).  roots = w8192^i (i = 0..8192).
// inverse DIT: bit-reversed input -> natural output (unscaled); twiddle w4096^(-j * 4096/m) = roots[8192 - 2*j*4096/m]
__device__ __forceinline__ void ntt4096_dit_inverse(Fr* sh, const Fr* __restrict__ roots, int tid) {
#pragma unroll 1
    for (int half = 1; half < NTT_N; half <<= 1) {  // half = m/2
        const int tw_step = (N_EXT / 2) / half;     // 8192 / m
#pragma unroll 1
        for (int b = tid; b < NTT_N / 2; b += NTT_THREADS) {
            int j = b & (half - 1);
            int i0 = ((b - j) << 1) + j, i1 = i0 + half;
            Fr u = sh[i0], v = sh[i1];
            if (j != 0) v = mul(v, ld_fr(roots + (N_EXT - j * tw_step)));
            sh[i0] = add(u, v);
            sh[i1] = sub(u, v);
        }
        __syncthreads();
    }
}
// forward DIF: natural input -> bit-reversed output; twiddle w4096^(j * 4096/m) = roots[2*j*4096/m]
__device__ __forceinline__ void ntt4096_dif_forward(Fr* sh, const Fr* __restrict__ roots, int tid) {
#pragma unroll 1
    for (int half = NTT_N / 2; half >= 1; half >>= 1) {
        const int tw_step = (N_EXT / 2) / half;
#pragma unroll 1
        for (int b = tid; b < NTT_N / 2; b += NTT_THREADS) {
            int j = b & (half - 1);
            int i0 = ((b - j) << 1) + j, i1 = i0 + half;
            Fr u = sh[i0], v = sh[i1];
            Fr d = sub(u, v);
            if (j != 0) d = mul(d, ld_fr(roots + j * tw_step));
            sh[i0] = add(u, v);
            sh[i1] = d;
        }
        __syncthreads();
    }
}

// blob -> monomial coefficients (global, Montgomery) and/or cells (bytes).
// TMA: the 128 KiB of the blob arrive in shared memory through ONE bulk asynchronous copy (cp.async.bulk + mbarrier,
// tma.cuh) into the buffer the transforms work in; every thread then converts its elements in place (raw big-endian
// bytes -> Montgomery form, same 32-byte slot).  Otherwise: two LDG.128 per element through registers.
template <bool TMA>
__global__ void __launch_bounds__(NTT_THREADS) blob_to_cells_kernel(uint8_t* __restrict__ cells, Fr* __restrict__ mono, const uint8_t* __restrict__ blobs,
                                                                    const Fr* __restrict__ roots, int* __restrict__ bad) {
    extern __shared__ uint4 smem_raw[];
    Fr* sh = reinterpret_cast<Fr*>(smem_raw);
    __shared__ int s_bad;
    __shared__ __align__(8) uint64_t s_bar;
    const int blob = blockIdx.x, tid = threadIdx.x;
    const uint8_t* src = blobs + (size_t)blob * BLOB_BYTES;
    if (tid == 0) {
        s_bad = 0;
        if (TMA) tma_mbar_init(&s_bar, 1);
    }
    __syncthreads();
    if (TMA) {
        if (tid == 0) {
            tma_mbar_expect_tx(&s_bar, BLOB_BYTES);
            tma_load_1d(smem_raw, src, BLOB_BYTES, &s_bar);
        }
        tma_mbar_wait(&s_bar, 0);
    }
    for (int i = tid; i < NTT_N; i += NTT_THREADS) {
        uint4 hi, lo;
        if (TMA) {
            hi = smem_raw[2 * i];
            lo = smem_raw[2 * i + 1];
        } else {
            const uint4* q = reinterpret_cast<const uint4*>(src + 32 * i);
            hi = __ldg(q);
            lo = __ldg(q + 1);
        }
        uint32_t s[8] = {bswap32c(lo.w), bswap32c(lo.z), bswap32c(lo.y), bswap32c(lo.x), bswap32c(hi.w), bswap32c(hi.z), bswap32c(hi.y), bswap32c(hi.x)};
        if (limbs_geq<8>(s, FR_MOD)) s_bad = 1;  // bytes_to_bls_field, src/common/bytes.c:67
        sh[i] = to_mont<FrTag>(s);
        if (cells) {  // cells 0..63 are the blob itself (bit-reversed extended evaluations, first half)
            uint4* d = reinterpret_cast<uint4*>(cells + (size_t)blob * 2 * BLOB_BYTES + 32 * i);
            d[0] = hi;
            d[1] = lo;
        }
    }
    __syncthreads();
    ntt4096_dit_inverse(sh, roots, tid);
    const Fr inv_n = Fr::from_limbs(FR_INV_4096);
    for (int i = tid; i < NTT_N; i += NTT_THREADS) {
        Fr a = mul(sh[i], inv_n);
        if (mono) mono[(size_t)blob * NTT_N + i] = a;
        sh[i] = mul(a, ld_fr(roots + i));  // a[m] * w8192^m : odd-index extended evaluations
    }
    __syncthreads();
    if (cells) {
        ntt4096_dif_forward(sh, roots, tid);
        for (int i = tid; i < NTT_N; i += NTT_THREADS) {
            uint32_t t[8];
            from_mont<FrTag>(t, sh[i]);
            uint4* d = reinterpret_cast<uint4*>(cells + (size_t)blob * 2 * BLOB_BYTES + BLOB_BYTES + 32 * i);
            d[0] = make_uint4(bswap32c(t[7]), bswap32c(t[6]), bswap32c(t[5]), bswap32c(t[4]));
            d[1] = make_uint4(bswap32c(t[3]), bswap32c(t[2]), bswap32c(t[1]), bswap32c(t[0]));
        }
    }
    if (tid == 0 && s_bad && bad) bad[blob] = 1;
}

// ------------------------------------------------------------------------------------------------
// FK20 phase 1: for every offset i < 64 the circulant vector c_i (fk20.c:55-78) and its FFT-128,
// pre-scaled by 1/128, stored as plain scalars S[blob][j][i] (the 64 scalars of MSM j contiguous).
// One CTA per (blob, group of 16 offsets): 16 x 128 Fr = 64 KiB of shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int FKS_THREADS = 256;
constexpr int FKS_GROUP = 16;

__global__ void __launch_bounds__(FKS_THREADS) fk20_scalars_kernel(uint32_t* __restrict__ S, const Fr* __restrict__ mono, const Fr* __restrict__ roots) {
    extern __shared__ uint4 smem_raw[];
    Fr* sh = reinterpret_cast<Fr*>(smem_raw);  // [16][128]
    const int blob = blockIdx.y, grp = blockIdx.x, tid = threadIdx.x;
    const Fr* p = mono + (size_t)blob * NTT_N;
    // fill: out[0] = in[d - i]; out[1..65] = 0; out[128 - j] = in[d - i - 64 j], j = 1..62   (d = 4095)
    for (int e = tid; e < FKS_GROUP * 128; e += FKS_THREADS) {
        int o = e >> 7, k = e & 127;
        int off = grp * FKS_GROUP + o;
        int dmi = 4095 - off;
        Fr v = Fr::zero();
        if (k == 0)
            v = ld_fr(p + dmi);
        else if (k >= 66)
            v = ld_fr(p + (dmi - 64 * (128 - k)));
        sh[e] = v;
    }
    __syncthreads();
    // forward DIF over each 128-vector; twiddle w128^(j * 128/m) = roots[64 * j * 128/m]
#pragma unroll 1
    for (int half = 64; half >= 1; half >>= 1) {
        const int tw_step = (N_EXT / 2) / half;
#pragma unroll 1
        for (int b = tid; b < FKS_GROUP * 64; b += FKS_THREADS) {
            int o = b >> 6, bb = b & 63;
            int j = bb & (half - 1);
            int i0 = (o << 7) + ((bb - j) << 1) + j, i1 = i0 + half;
            Fr u = sh[i0], v = sh[i1];
            Fr d = sub(u, v);
            if (j != 0) d = mul(d, ld_fr(roots + j * tw_step));
            sh[i0] = add(u, v);
            sh[i1] = d;
        }
        __syncthreads();
    }
    // position q holds frequency brp7(q); scale by 1/128 (fk20.c:188-209), leave Montgomery form
    const Fr inv128 = Fr::from_limbs(FR_INV_128);
    for (int e = tid; e < FKS_GROUP * 128; e += FKS_THREADS) {
        int o = e >> 7, q = e & 127;
        int j = (int)(__brev((uint32_t)q) >> 25);
        int off = grp * FKS_GROUP + o;
        uint32_t t[8];
        from_mont<FrTag>(t, mul(sh[e], inv128));
        uint4* d = reinterpret_cast<uint4*>(S + (((size_t)blob * 128 + j) * 64 + off) * 8);
        d[0] = make_uint4(t[0], t[1], t[2], t[3]);
        d[1] = make_uint4(t[4], t[5], t[6], t[7]);
    }
}

int launch_blob_to_cells(Launch& L, uint8_t* cells, Fr* mono, const uint8_t* blobs, uint64_t n, int* d_bad) {
    if (!n) return RET_OK;
    // CKZG_B200_CELLS_LOAD=ldg: register loads instead of the bulk copy (A/B measurements, profiles/)
    static const bool use_tma = !(getenv("CKZG_B200_CELLS_LOAD") && strcmp(getenv("CKZG_B200_CELLS_LOAD"), "ldg") == 0);
    if (use_tma) {
        KZG_FUNC_ATTR_PER_DEVICE(blob_to_cells_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_N * (int)sizeof(Fr));
        blob_to_cells_kernel<true><<<(unsigned)n, NTT_THREADS, NTT_N * sizeof(Fr), L.stream>>>(cells, mono, blobs, L.ctx->roots, d_bad);
    } else {
        KZG_FUNC_ATTR_PER_DEVICE(blob_to_cells_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_N * (int)sizeof(Fr));
        blob_to_cells_kernel<false><<<(unsigned)n, NTT_THREADS, NTT_N * sizeof(Fr), L.stream>>>(cells, mono, blobs, L.ctx->roots, d_bad);
    }
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "blob_to_cells");
    return RET_OK;
}

int launch_fk20_scalars(Launch& L, uint32_t* S, const Fr* mono, uint64_t n) {
    if (!n) return RET_OK;
    KZG_FUNC_ATTR_PER_DEVICE(fk20_scalars_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FKS_GROUP * 128 * (int)sizeof(Fr));
    dim3 grid(64 / FKS_GROUP, (unsigned)n);
    fk20_scalars_kernel<<<grid, FKS_THREADS, FKS_GROUP * 128 * sizeof(Fr), L.stream>>>(S, mono, L.ctx->roots);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "fk20_scalars");
    return RET_OK;
}

}  // namespace kzg
