// G1 FFTs of FK20 (128 points per vector) with every group operation spread over four lanes.
//
// Replaces (paths relative to the reference tree):
//   g1_fft_fast / g1_fft / g1_ifft_unscaled ... src/eip7594/fft.c:164-240
//   their use in compute_fk20_cell_proofs ..... src/eip7594/fk20.c:257-269
//   and in init_fk20_multi_settings ........... src/setup/setup.c:284-289
//
// A radix-2 stage is 64 twiddle multiplications per vector, and the 14 stages of the two transforms are a
// dependency chain: with one multiplication per thread (129 doublings + ~52 additions, ~2000 dependent Fp
// products, 1.9 ms) a batch of 256 blobs put one warp on each SM sub-partition and waited 14 times for
// that chain (r01y: 43 ms of the 71 ms per 256 blobs).  Here a QUAD of lanes shares each group operation
// (g1_quad.cuh: doubling = 3 product levels instead of 9 products, addition = 4 instead of 14), which cuts
// the chain to ~630 product latencies and puts four times as many warps on the multiplier.
//
// Thread layout: CTA = 4 warps = 4 butterflies; the 8 quads of a warp take the SAME butterfly of 8
// different vectors, so the GLV + width-4-NAF digit strings of the twiddle (fft_twiddles.cuh) are
// warp-uniform.  Per quad, shared memory holds the table of odd multiples, the beta-twisted x
// coordinates (second GLV base -phi(P) = (beta x, -y)), the accumulator and the scratch of the
// cooperative operations.
#define KZG_FP_MUL_OUTLINE 1
#include "cells.h"
#include "fft_twiddles.cuh"
#include "g1_quad.cuh"

namespace kzg {

struct QuadWork {  // shared memory, per quad
    G1 acc;
    G1 t;
    QuadScratch sc;
};
struct QuadTab {  // global scratch (L2), per quad: read once per addition, so its latency hardly shows,
    G1 tab[8];    // (2i+1) P                    and leaving it out of shared memory lets a whole 256-blob
    Fp bx[8];     // beta * tab[i].x             stage be resident at once (7 CTAs per SM instead of 2)
};
constexpr int GQ_WARPS = 4;
constexpr int GQ_QUADS = GQ_WARPS * 8;
constexpr size_t GQ_SMEM = GQ_QUADS * sizeof(QuadWork);

// W->acc = [w128^e] *src   (e != 0; src may be W->t)
static __device__ __noinline__ void g1_mul_twiddle_quad(QuadWork* W, QuadTab* T, const G1* src, int e) {
    const unsigned lane = threadIdx.x & 31u, ql = lane & 3u;
    const unsigned mask = 0xFu << (lane & ~3u);
    quad_copy_g1(&T->tab[0], src);
    __syncwarp(mask);
    g1_dbl_quad(&W->t, &T->tab[0], &W->sc);
#pragma unroll 1
    for (int i = 1; i < 8; i++) g1_add_quad(&T->tab[i], &T->tab[i - 1], &W->t, &W->sc);
    {
        const Fp beta = Fp::from_limbs(FP_BETA_A);
#pragma unroll 1
        for (int i = (int)ql; i < 8; i += 4) quad_st(&T->bx[i], mul(quad_ld(&T->tab[i].x), beta));
        if (ql == 0) {
            const Fp z = Fp::zero();
            quad_st(&W->acc.x, z); quad_st(&W->acc.y, z); quad_st(&W->acc.zz, z); quad_st(&W->acc.zzz, z);
        }
    }
    __syncwarp(mask);
    const int8_t* d1 = FFT_TW_NAF[e][0];
    const int8_t* d2 = FFT_TW_NAF[e][1];
#pragma unroll 1
    for (int i = FFT_TW_TOP[e] - 1; i >= 0; i--) {
        g1_dbl_quad(&W->acc, &W->acc, &W->sc);
        const int a = d1[i], b = d2[i];
        if (a != 0) {
            const G1* t = &T->tab[((a < 0 ? -a : a) - 1) >> 1];
            g1_add_quad_q(&W->acc, &W->acc, &t->x, t, a < 0, &W->sc);
        }
        if (b != 0) {  // base -phi(P): (beta x, -y)
            const int idx = ((b < 0 ? -b : b) - 1) >> 1;
            g1_add_quad_q(&W->acc, &W->acc, &T->bx[idx], &T->tab[idx], b > 0, &W->sc);
        }
    }
}

enum { GS_INVERSE = 0, GS_FORWARD_FIRST = 1, GS_FORWARD = 2 };

// One radix-2 stage, data in global memory (L2).
// mode GS_INVERSE (decimation in time, g1_ifft_unscaled fft.c:227): v' = [w^-j] v; (u+v', u-v'); the
//   last stage (half = 64) keeps only the lower output (FK20 discards the upper half, fk20.c:264-266).
// mode GS_FORWARD_FIRST: input upper half is infinity: (u, [w^j] u).
// mode GS_FORWARD (decimation in frequency, g1_fft fft.c:199): (u+v, [w^j](u-v)).
__global__ void __launch_bounds__(32 * GQ_WARPS, 4) g1_fft_stage_quad_kernel(G1* __restrict__ data, QuadTab* __restrict__ tabs, uint64_t nvec, int half, int mode) {
    extern __shared__ __align__(16) unsigned char gq_smem[];
    QuadWork* W = reinterpret_cast<QuadWork*>(gq_smem) + (threadIdx.x >> 2);
    QuadTab* T = tabs + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * GQ_QUADS + (threadIdx.x >> 2));
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned mask = 0xFu << (lane & ~3u);
    const int b = blockIdx.x * GQ_WARPS + warp;  // butterfly 0..63
    const uint64_t vec = (uint64_t)blockIdx.y * 8 + (lane >> 2);
    if (vec >= nvec) return;
    const int j = b & (half - 1);
    const int i0 = ((b - j) << 1) + j;
    const int step = 64 / half;  // twiddle exponent of w128 per unit of j
    G1* P0 = data + vec * 128 + i0;
    G1* P1 = P0 + half;
    if (mode == GS_FORWARD_FIRST) {
        const G1* r = P0;
        if (j != 0) {
            g1_mul_twiddle_quad(W, T, P0, j * step);
            r = &W->acc;
        }
        quad_copy_g1(P1, r);
        return;
    }
    if (mode == GS_INVERSE) {
        const G1* v = P1;
        if (j != 0) {
            g1_mul_twiddle_quad(W, T, P1, (128 - j * step) & 127);
            v = &W->acc;
        }
        if (half != 64) {
            g1_add_quad_q(&W->t, P0, &v->x, v, true, &W->sc);  // u - v'
            g1_add_quad(P0, P0, v, &W->sc);                    // u + v'
            quad_copy_g1(P1, &W->t);
        } else {
            g1_add_quad(P0, P0, v, &W->sc);
        }
        return;
    }
    g1_add_quad_q(&W->t, P0, &P1->x, P1, true, &W->sc);  // u - v
    g1_add_quad(P0, P0, P1, &W->sc);                     // u + v
    const G1* r = &W->t;
    if (j != 0) {
        g1_mul_twiddle_quad(W, T, &W->t, j * step);
        r = &W->acc;
    }
    __syncwarp(mask);
    quad_copy_g1(P1, r);
}

// ------------------------------------------------------------------------------------------------
// Large batches (>= 1024 vectors per call): ONE THREAD per butterfly.  With 64 x nvec butterflies per stage there are
// enough independent chains to fill every sub-partition with several warps, and a thread's XYZZ operations use every
// multiplier slot they occupy (the quad form spends 12 slots on the 9 products of a doubling and 16 on the 14 of an
// addition, and idles a quarter of its lanes in the sums) -- the same trade fk20_msm_kernel makes against the affine
// levels.  All 128 threads of a CTA take the SAME butterfly of 128 different vectors, so the width-4 NAF digit strings
// of the twiddle are CTA-uniform and nothing diverges.  The table of odd multiples lives in a global scratch (L2):
// it is read once per addition, i.e. every ~2.5 doublings.
// ------------------------------------------------------------------------------------------------
struct ThreadTab {
    G1 tab[8];  // (2i+1) P
    Fp bx[8];   // beta * tab[i].x  (second GLV base -phi(P) = (beta x, -y))
};
constexpr int GT_THREADS = 128;

static __device__ __noinline__ G1 g1_mul_twiddle_thread(ThreadTab* __restrict__ T, const G1& src, int e) {
    if (g1_is_inf(src)) return g1_inf();
    G1 p = src, d = src;
    g1_store(&T->tab[0], p);
    g1_dbl_to(d);
#pragma unroll 1
    for (int i = 1; i < 8; i++) {
        g1_add_to(p, d);
        g1_store(&T->tab[i], p);
    }
    {
        const Fp beta = Fp::from_limbs(FP_BETA_A);
        p = src;
        T->bx[0] = mul(p.x, beta);
#pragma unroll 1
        for (int i = 1; i < 8; i++) T->bx[i] = mul(g1_load(&T->tab[i]).x, beta);
    }
    G1 acc = g1_inf();
    const int8_t* d1 = FFT_TW_NAF[e][0];
    const int8_t* d2 = FFT_TW_NAF[e][1];
#pragma unroll 1
    for (int i = FFT_TW_TOP[e] - 1; i >= 0; i--) {
        g1_dbl_to(acc);
        const int a = d1[i], b = d2[i];
        if (a != 0) {
            G1 t = g1_load(&T->tab[((a < 0 ? -a : a) - 1) >> 1]);
            if (a < 0) t.y = neg(t.y);
            g1_add_to(acc, t);
        }
        if (b != 0) {
            const int idx = ((b < 0 ? -b : b) - 1) >> 1;
            G1 t = g1_load(&T->tab[idx]);
            t.x = T->bx[idx];
            if (b > 0) t.y = neg(t.y);
            g1_add_to(acc, t);
        }
    }
    return acc;
}

// same stage semantics as g1_fft_stage_quad_kernel; grid = (64 butterflies, ceil(nvec / 128))
__global__ void __launch_bounds__(GT_THREADS, 2) g1_fft_stage_thread_kernel(G1* __restrict__ data, ThreadTab* __restrict__ tabs, uint64_t nvec, int half, int mode) {
    const int b = blockIdx.x;
    const uint64_t vec = (uint64_t)blockIdx.y * GT_THREADS + threadIdx.x;
    if (vec >= nvec) return;
    ThreadTab* T = tabs + ((size_t)blockIdx.y * 64 + b) * GT_THREADS + threadIdx.x;
    const int j = b & (half - 1);
    const int i0 = ((b - j) << 1) + j;
    const int step = 64 / half;
    G1* P0 = data + vec * 128 + i0;
    G1* P1 = P0 + half;
    if (mode == GS_FORWARD_FIRST) {
        G1 u = g1_load(P0);
        if (j != 0) u = g1_mul_twiddle_thread(T, u, j * step);
        g1_store(P1, u);
        return;
    }
    if (mode == GS_INVERSE) {
        G1 v = g1_load(P1);
        if (j != 0) v = g1_mul_twiddle_thread(T, v, (128 - j * step) & 127);
        G1 u = g1_load(P0);
        if (half != 64) {
            G1 m = v;
            m.y = neg(m.y);
            G1 dlt = u;
            g1_add_to(dlt, m);  // u - v'
            g1_store(P1, dlt);
        }
        g1_add_to(u, v);  // u + v'
        g1_store(P0, u);
        return;
    }
    G1 u = g1_load(P0), v = g1_load(P1);
    G1 m = v;
    m.y = neg(m.y);
    G1 dlt = u;
    g1_add_to(dlt, m);  // u - v
    g1_add_to(u, v);    // u + v
    g1_store(P0, u);
    if (j != 0) dlt = g1_mul_twiddle_thread(T, dlt, j * step);
    g1_store(P1, dlt);
}

static int g1_fft128_run_threads(Launch& L, G1* data, uint64_t nvec, bool with_inverse) {
    dim3 grid(64, (unsigned)((nvec + GT_THREADS - 1) / GT_THREADS));
    ThreadTab* tabs = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&tabs, (size_t)grid.x * grid.y * GT_THREADS * sizeof(ThreadTab), L.stream));
    if (with_inverse) {
        for (int half = 1; half <= 64; half <<= 1) {
            g1_fft_stage_thread_kernel<<<grid, GT_THREADS, 0, L.stream>>>(data, tabs, nvec, half, GS_INVERSE);
            KZG_CUDA_TRY(cudaGetLastError());
        }
    }
    g1_fft_stage_thread_kernel<<<grid, GT_THREADS, 0, L.stream>>>(data, tabs, nvec, 64, GS_FORWARD_FIRST);
    KZG_CUDA_TRY(cudaGetLastError());
    for (int half = 32; half >= 1; half >>= 1) {
        g1_fft_stage_thread_kernel<<<grid, GT_THREADS, 0, L.stream>>>(data, tabs, nvec, half, GS_FORWARD);
        KZG_CUDA_TRY(cudaGetLastError());
    }
    KZG_CUDA_TRY(cudaFreeAsync(tabs, L.stream));
    return RET_OK;
}

// in place: [inverse DIT on bit-reversed input, lower half kept] -> forward DIF with upper half = infinity
int g1_fft128_run(Launch& L, G1* data, uint64_t nvec, bool with_inverse) {
    // CKZG_B200_FFT_THREAD_MIN: batch size from which one thread takes a butterfly (0 = never)
    // (read on every call, not cached: tests switch it inside one process)
    const char* tm_env = getenv("CKZG_B200_FFT_THREAD_MIN");
    const uint64_t thread_min = tm_env ? (uint64_t)atoll(tm_env) : 1024;
    if (thread_min && nvec >= thread_min) return g1_fft128_run_threads(L, data, nvec, with_inverse);
    KZG_FUNC_ATTR_PER_DEVICE(g1_fft_stage_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GQ_SMEM);
    dim3 grid(64 / GQ_WARPS, (unsigned)((nvec + 7) / 8));
    QuadTab* tabs = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&tabs, (size_t)grid.x * grid.y * GQ_QUADS * sizeof(QuadTab), L.stream));
    if (with_inverse) {
        for (int half = 1; half <= 64; half <<= 1) {
            g1_fft_stage_quad_kernel<<<grid, 32 * GQ_WARPS, GQ_SMEM, L.stream>>>(data, tabs, nvec, half, GS_INVERSE);
            KZG_CUDA_TRY(cudaGetLastError());
        }
    }
    g1_fft_stage_quad_kernel<<<grid, 32 * GQ_WARPS, GQ_SMEM, L.stream>>>(data, tabs, nvec, 64, GS_FORWARD_FIRST);
    KZG_CUDA_TRY(cudaGetLastError());
    for (int half = 32; half >= 1; half >>= 1) {
        g1_fft_stage_quad_kernel<<<grid, 32 * GQ_WARPS, GQ_SMEM, L.stream>>>(data, tabs, nvec, half, GS_FORWARD);
        KZG_CUDA_TRY(cudaGetLastError());
    }
    KZG_CUDA_TRY(cudaFreeAsync(tabs, L.stream));
    return RET_OK;
}

}  // namespace kzg
