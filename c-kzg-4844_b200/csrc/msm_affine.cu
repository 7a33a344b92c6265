// Batched-affine form of the bucket-free commitment MSM (msm_direct.cu) for batches of blobs.
//
// Same sum as msm_direct_kernel -- the 4096 x W table points D[i][w][digit] selected by a blob's scalars --
// but added PAIRWISE IN AFFINE COORDINATES, level by level, with the inversions batched (Montgomery's
// trick): an affine addition is lambda = (y2 - y1) / (x2 - x1), x3 = lambda^2 - x1 - x2,
// y3 = lambda (x1 - x3) - y1, and with the denominators of a thread's 19..32 pairs inverted together
// (one prefix product going up, two products coming down, ONE field inversion per CTA of segments) it costs
// 6 Fp products against the 10 of an XYZZ mixed addition.  The inversion itself (binary Euclid, field.cuh)
// runs on the ALU pipe, which the multiplier-bound kernels leave idle.
//
//   level 1  (bam_level1_kernel): thread t of a blob pairs the table points of scalars (t, t + 1024) and of
//            (t + 2048, t + 3072), window by window: 2 segments of W pairs -> 2048 W affine points per blob;
//   levels 2.. (bam_level_kernel): pairs (2j, 2j + 1) of the previous level, 32 pairs per thread, until
//            ~1.2 k points per blob are left;
//   final    (bam_final_kernel): one CTA per blob adds what is left into an XYZZ accumulator + tree.
// Exceptional pairs are exact, not assumed away: infinity operands (zero digits) pass the other operand
// through, equal points take the doubling slope 3 x^2 / (2 y) through the same batched inversion, opposite
// points give infinity.
//
// Replaces g1_lincomb_fast (src/common/lincomb.c:65) for the fixed bases of blob_to_kzg_commitment and of the
// quotient commitment, like msm_direct.cu -- behind CKZG_B200_AFFINE_MIN (off by default: with the 61 GB table
// level 1 is bound by its two passes of random HBM reads and the form only reaches parity, r02p) -- and, in
// the second half of this file, the 128 x MSM(64) of FK20 for batches of >= 8 blobs, where it is 20 % faster
// than the XYZZ kernel (r02q).  A single blob stays on the one-kernel XYZZ forms: every level here is a
// latency chain of its own.
#define KZG_FP_MUL_OUTLINE 1
#include "affine_batch.cuh"
#include "cells.h"

namespace kzg {

constexpr int BAM_THREADS = 128;
constexpr int BAM_SEG = 32;         // pairs per thread in the generic levels
constexpr int BAM_STOP = 1536;      // stop halving once a blob has at most this many points

__device__ __forceinline__ uint32_t bam_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ Fp bam_ld_fp(const Fp* p) {
    Fp a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __ldg(q + i);
    return a;
}
__device__ __forceinline__ void bam_st_fp(Fp* p, const Fp& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* d = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = d[i];
}

// 1 / run for every thread of the CTA with ONE inversion: all lanes of one warp invert the same value -- the
// product of all segment products -- so the branchy binary-Euclid loop runs without divergence (with 32
// different inputs a warp walks all four branch bodies on almost every step: measured, the per-lane
// inversions doubled the cost of a level).  Exclusive prefix and suffix products by two shuffle scans:
// 1 / run_l = (1 / total) * prefix_l * suffix_l.  run must be non-zero on every lane.
__device__ __forceinline__ Fp bam_shfl_up(const Fp& v, int delta) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], delta);
    return r;
}
__device__ __forceinline__ Fp bam_shfl_down(const Fp& v, int delta) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], delta);
    return r;
}
__device__ __forceinline__ Fp bam_shfl(const Fp& v, int src) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}
struct BamShare {
    Fp total[BAM_THREADS / 32];
    Fp inv[BAM_THREADS / 32];
};
// Called by ALL threads of the CTA.  One inversion per CTA: warp 0 inverts the product of the four warp totals
// while the other warps wait at the barrier (their sub-partitions serve the other CTAs of the SM meanwhile).
static __device__ __noinline__ Fp bam_cta_inverse(const Fp run, BamShare* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Fp inc = run, dec = run;  // inclusive prefix / suffix products inside the warp
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const Fp up = bam_shfl_up(inc, d), dn = bam_shfl_down(dec, d);
        const Fp a = mul(inc, up), b = mul(dec, dn);
        if (lane >= d) inc = a;
        if (lane + d < 32) dec = b;
    }
    if (lane == 31) sh->total[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = BAM_THREADS / 32;
        Fp all = sh->total[0];
#pragma unroll
        for (int v = 1; v < NW; v++) all = mul(all, sh->total[v]);
        const Fp ainv = fp_inv(all);  // the same value on every lane: no divergence
        if (lane < NW) {
            Fp others = Fp::one();
            for (int v = 0; v < NW; v++)
                if (v != lane) others = mul(others, sh->total[v]);
            sh->inv[lane] = mul(ainv, others);
        }
    }
    __syncthreads();
    const Fp tinv = sh->inv[warp];
    // exclusive products: shift the inclusive ones by one lane
    Fp pre = bam_shfl_up(inc, 1), suf = bam_shfl_down(dec, 1);
    if (lane == 0) pre = Fp::one();
    if (lane == 31) suf = Fp::one();
    const Fp r = mul(tinv, mul(pre, suf));
    __syncthreads();  // sh is reused by the next call
    return r;
}

// signed c-bit digits of a scalar (magnitude | sign << 15), as msm_direct.cu
__device__ __forceinline__ void bam_digits(uint16_t* dig, const uint8_t* scalars, uint64_t elem, bool big_endian, const FkGeom& g, int* bad, uint64_t blob) {
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + elem * 32);
    const uint4 a = __ldg(sp), b = __ldg(sp + 1);
    uint32_t s[9];
    bool ok = true;
    if (big_endian) {  // wire form, must be canonical (bytes_to_bls_field, src/common/bytes.c:64)
        s[0] = bam_bswap(b.w); s[1] = bam_bswap(b.z); s[2] = bam_bswap(b.y); s[3] = bam_bswap(b.x);
        s[4] = bam_bswap(a.w); s[5] = bam_bswap(a.z); s[6] = bam_bswap(a.y); s[7] = bam_bswap(a.x);
        if (limbs_geq<8>(s, FR_MOD)) {
            if (bad) bad[blob] = 1;
            ok = false;
        }
    } else {
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
        s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    }
    s[8] = 0;
    const uint32_t dmask = (1u << g.c) - 1u, dfull = 1u << g.c;
    uint32_t carry = 0;
#pragma unroll 1
    for (int w = 0; w < g.w; w++) {
        const int o = w * g.c;
        uint32_t d = (__funnelshift_r(s[o >> 5], s[(o >> 5) + 1], o & 31) & dmask) + carry;
        const bool negd = d > (uint32_t)g.m;
        carry = negd ? 1u : 0u;
        const uint32_t mag = negd ? (dfull - d) : d;
        dig[w] = ok ? (uint16_t)(mag | (negd ? 0x8000u : 0u)) : (uint16_t)0;
    }
}

// ---- level 1: table points of two pairs of scalars, window by window -----------------------------------
constexpr int BAM_L1_MAX = 2 * 22;  // pairs per thread: 2 scalar pairs x W windows (W <= 22 for c >= 12)
__global__ void __launch_bounds__(BAM_THREADS, 4) bam_level1_kernel(G1Affine* __restrict__ out, const uint8_t* __restrict__ scalars, bool big_endian,
                                                                    const G1Affine* __restrict__ table, int* __restrict__ bad, const FkGeom g) {
    __shared__ BamShare share;
    const int t = blockIdx.x * BAM_THREADS + threadIdx.x;  // 0..1023
    const uint64_t blob = blockIdx.y;
    G1Affine* dst = out + blob * (size_t)(2048 * g.w);
    uint16_t dig[4][22];  // scalars t, t + 1024, t + 2048, t + 3072
    Fp prefix[BAM_L1_MAX];
    const int W = g.w, np = 2 * g.w;
#pragma unroll 1
    for (int q = 0; q < 4; q++) bam_digits(dig[q], scalars, blob * N_BLOB + t + 1024 * q, big_endian, g, bad, blob);
    // pair k = seg * W + w adds window w of scalar 2 seg and of scalar 2 seg + 1
    auto operand = [&](int k, int side, bool& inf, bool& negy) -> const G1Affine* {
        const int seg = k >= W ? 1 : 0, w = k - seg * W, q = 2 * seg + side;
        const uint16_t dg = dig[q][w];
        inf = (dg & 0x7fffu) == 0;
        negy = (dg & 0x8000u) != 0;
        const G1Affine* tp = table + ((size_t)(t + 1024 * q) * W) * g.m;
        return inf ? tp : tp + (size_t)w * g.m + ((dg & 0x7fffu) - 1u);
    };
    // the gathers are random 96-byte reads over a table of up to 61 GB: ask for all of them first
#pragma unroll 1
    for (int k = 0; k < np; k++) {
        bool ia, ib, na, nb;
        const char* ea = reinterpret_cast<const char*>(operand(k, 0, ia, na));
        const char* eb = reinterpret_cast<const char*>(operand(k, 1, ib, nb));
        if (!ia) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ea));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ea + 64));
        }
        if (!ib) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(eb));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(eb + 64));
        }
    }
    // pass 1: denominators and their running product
    Fp run = Fp::one();
#pragma unroll 1
    for (int k = 0; k < np; k++) {
        bool inf_a, inf_b, na, nb;
        const G1Affine* pa = operand(k, 0, inf_a, na);
        const G1Affine* pb = operand(k, 1, inf_b, nb);
        const Fp xa = bam_ld_fp(&pa->x), xb = bam_ld_fp(&pb->x);
        // a table entry can itself be infinity, stored as (0, 0) (a setup point at infinity): x = 0 is not on the curve otherwise
        inf_a = inf_a || (is_zero(xa) && is_zero(bam_ld_fp(&pa->y)));
        inf_b = inf_b || (is_zero(xb) && is_zero(bam_ld_fp(&pb->y)));
        Fp d;
        (void)bam_plan(d, xa, xb, inf_a, inf_b, [&] { return cneg(bam_ld_fp(&pa->y), na); }, [&] { return cneg(bam_ld_fp(&pb->y), nb); });
        prefix[k] = run;
        run = mul(run, d);
    }
    Fp inv = bam_cta_inverse(run, &share);
    // pass 2, backwards: 1 / d_k = inv * prefix[k]; inv *= d_k
#pragma unroll 1
    for (int k = np - 1; k >= 0; k--) {
        bool inf_a, inf_b, na, nb;
        const G1Affine* pa = operand(k, 0, inf_a, na);
        const G1Affine* pb = operand(k, 1, inf_b, nb);
        const Fp xa = bam_ld_fp(&pa->x), xb = bam_ld_fp(&pb->x);
        const Fp ya = cneg(bam_ld_fp(&pa->y), na), yb = cneg(bam_ld_fp(&pb->y), nb);
        inf_a = inf_a || (is_zero(xa) && is_zero(ya));
        inf_b = inf_b || (is_zero(xb) && is_zero(yb));
        Fp d;
        const int kind = bam_plan(d, xa, xb, inf_a, inf_b, [&] { return ya; }, [&] { return yb; });
        const Fp dinv = mul(inv, prefix[k]);
        inv = mul(inv, d);
        Fp x3, y3;
        bam_finish(x3, y3, kind, xa, ya, xb, yb, dinv);
        const int seg = k >= W ? 1 : 0, w = k - seg * W;
        G1Affine* o = dst + ((size_t)(seg * 1024 + t) * W + w);
        bam_st_fp(&o->x, x3);
        bam_st_fp(&o->y, y3);
    }
}

// ---- levels 2..: out[j] = in[2j] + in[2j+1], `half` pairs per blob, BAM_SEG pairs per thread -----------
__global__ void __launch_bounds__(BAM_THREADS, 4) bam_level_kernel(G1Affine* __restrict__ out, const G1Affine* __restrict__ in, uint32_t half) {
    __shared__ BamShare share;
    const uint64_t blob = blockIdx.y;
    const uint32_t j0 = (blockIdx.x * BAM_THREADS + threadIdx.x) * BAM_SEG;
    // lanes past the end stay for the warp-wide inversion with an empty segment
    const uint32_t cnt = j0 >= half ? 0u : (half - j0 < (uint32_t)BAM_SEG ? half - j0 : (uint32_t)BAM_SEG);
    const G1Affine* src = in + blob * (size_t)(2 * half) + 2 * (size_t)j0;
    G1Affine* dst = out + blob * (size_t)half + j0;
    Fp prefix[BAM_SEG];
    Fp run = Fp::one();
#pragma unroll 1
    for (uint32_t k = 0; k < cnt; k++) {
        const G1Affine* pa = src + 2 * k;
        const G1Affine* pb = pa + 1;
        const Fp xa = bam_ld_fp(&pa->x), xb = bam_ld_fp(&pb->x);
        // infinity is (0, 0); x = 0 alone is not a point of the curve's subgroup either way (y^2 = 4 has y = +-2 of order 3)
        const bool inf_a = is_zero(xa) && is_zero(bam_ld_fp(&pa->y)), inf_b = is_zero(xb) && is_zero(bam_ld_fp(&pb->y));
        Fp d;
        (void)bam_plan(d, xa, xb, inf_a, inf_b, [&] { return bam_ld_fp(&pa->y); }, [&] { return bam_ld_fp(&pb->y); });
        prefix[k] = run;
        run = mul(run, d);
    }
    Fp inv = bam_cta_inverse(run, &share);
#pragma unroll 1
    for (int k = (int)cnt - 1; k >= 0; k--) {
        const G1Affine* pa = src + 2 * k;
        const G1Affine* pb = pa + 1;
        const Fp xa = bam_ld_fp(&pa->x), xb = bam_ld_fp(&pb->x), ya = bam_ld_fp(&pa->y), yb = bam_ld_fp(&pb->y);
        const bool inf_a = is_zero(xa) && is_zero(ya), inf_b = is_zero(xb) && is_zero(yb);
        Fp d;
        const int kind = bam_plan(d, xa, xb, inf_a, inf_b, [&] { return ya; }, [&] { return yb; });
        const Fp dinv = mul(inv, prefix[k]);
        inv = mul(inv, d);
        Fp x3, y3;
        bam_finish(x3, y3, kind, xa, ya, xb, yb, dinv);
        bam_st_fp(&dst[k].x, x3);
        bam_st_fp(&dst[k].y, y3);
    }
}

// ---- final: result[blob] = sum of the blob's `count` remaining affine points ---------------------------
__global__ void __launch_bounds__(BAM_THREADS) bam_final_kernel(G1* __restrict__ result, const G1Affine* __restrict__ in, uint32_t count) {
    __shared__ G1 sh[BAM_THREADS];
    const int t = threadIdx.x;
    const uint64_t blob = blockIdx.x;
    const G1Affine* src = in + blob * (size_t)count;
    G1 acc = g1_inf();
#pragma unroll 1
    for (uint32_t k = t; k < count; k += BAM_THREADS) {
        G1Affine a;
        a.x = bam_ld_fp(&src[k].x);
        a.y = bam_ld_fp(&src[k].y);
        g1_madd_to(acc, a, false);
    }
    sh[t] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = BAM_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) {
            G1 x = sh[t], y = sh[t + s];
            g1_add_to(x, y);
            sh[t] = x;
        }
        __syncthreads();
    }
    if (t < 12) reinterpret_cast<uint4*>(result + blob)[t] = reinterpret_cast<const uint4*>(&sh[0])[t];
}

// ------------------------------------------------------------------------------------------------
// The same scheme for the 128 x MSM(64) of FK20 (fk20.cu fk20_msm_kernel does them with XYZZ additions)
// ------------------------------------------------------------------------------------------------
// level 1: 16 threads per MSM; thread r pairs scalars (r, r + 32) and (r + 16, r + 48), window by window
__global__ void __launch_bounds__(BAM_THREADS, 4) bam_fk_level1_kernel(G1Affine* __restrict__ out, const uint32_t* __restrict__ S, const G1Affine* __restrict__ table,
                                                                       uint64_t total_msm, const FkGeom g) {
    __shared__ BamShare share;
    const uint64_t gid = (uint64_t)blockIdx.x * BAM_THREADS + threadIdx.x;
    const uint64_t msm = gid >> 4;  // = blob * 128 + j
    const int r = (int)(gid & 15);
    const bool live = msm < total_msm;
    const int j = (int)(msm & 127);
    const int W = g.w, np = live ? 2 * g.w : 0;
    uint16_t dig[4][32];
    Fp prefix[2 * 32];
    int idx[4];
#pragma unroll
    for (int q = 0; q < 4; q++) idx[q] = r + 16 * (q >> 1) + 32 * (q & 1);  // q = 2 seg + side
    if (live) {
#pragma unroll 1
        for (int q = 0; q < 4; q++) bam_digits(dig[q], reinterpret_cast<const uint8_t*>(S), msm * 64 + idx[q], false, g, nullptr, 0);
    }
    auto operand = [&](int k, int side, bool& inf, bool& negy) -> const G1Affine* {
        const int seg = k >= W ? 1 : 0, w = k - seg * W, q = 2 * seg + side;
        const uint16_t dg = dig[q][w];
        inf = (dg & 0x7fffu) == 0;
        negy = (dg & 0x8000u) != 0;
        const G1Affine* tp = table + ((size_t)(j * 64 + idx[q]) * W) * g.m;
        return inf ? tp : tp + (size_t)w * g.m + ((dg & 0x7fffu) - 1u);
    };
#pragma unroll 1
    for (int k = 0; k < np; k++) {
        bool ia, ib, na, nb;
        const char* ea = reinterpret_cast<const char*>(operand(k, 0, ia, na));
        const char* eb = reinterpret_cast<const char*>(operand(k, 1, ib, nb));
        if (!ia) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ea));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ea + 64));
        }
        if (!ib) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(eb));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(eb + 64));
        }
    }
    Fp run = Fp::one();
#pragma unroll 1
    for (int k = 0; k < np; k++) {
        bool inf_a, inf_b, na, nb;
        const G1Affine* pa = operand(k, 0, inf_a, na);
        const G1Affine* pb = operand(k, 1, inf_b, nb);
        const Fp xa = bam_ld_fp(&pa->x), xb = bam_ld_fp(&pb->x);
        inf_a = inf_a || (is_zero(xa) && is_zero(bam_ld_fp(&pa->y)));  // X^ columns do contain infinities (setup.c:272-282)
        inf_b = inf_b || (is_zero(xb) && is_zero(bam_ld_fp(&pb->y)));
        Fp d;
        (void)bam_plan(d, xa, xb, inf_a, inf_b, [&] { return cneg(bam_ld_fp(&pa->y), na); }, [&] { return cneg(bam_ld_fp(&pb->y), nb); });
        prefix[k] = run;
        run = mul(run, d);
    }
    Fp inv = bam_cta_inverse(run, &share);
    G1Affine* dst = out + msm * (size_t)(32 * W);
#pragma unroll 1
    for (int k = np - 1; k >= 0; k--) {
        bool inf_a, inf_b, na, nb;
        const G1Affine* pa = operand(k, 0, inf_a, na);
        const G1Affine* pb = operand(k, 1, inf_b, nb);
        const Fp xa = bam_ld_fp(&pa->x), xb = bam_ld_fp(&pb->x);
        const Fp ya = cneg(bam_ld_fp(&pa->y), na), yb = cneg(bam_ld_fp(&pb->y), nb);
        inf_a = inf_a || (is_zero(xa) && is_zero(ya));
        inf_b = inf_b || (is_zero(xb) && is_zero(yb));
        Fp d;
        const int kind = bam_plan(d, xa, xb, inf_a, inf_b, [&] { return ya; }, [&] { return yb; });
        const Fp dinv = mul(inv, prefix[k]);
        inv = mul(inv, d);
        Fp x3, y3;
        bam_finish(x3, y3, kind, xa, ya, xb, yb, dinv);
        const int seg = k >= W ? 1 : 0, w = k - seg * W;
        G1Affine* o = dst + ((size_t)(seg * 16 + r) * W + w);
        bam_st_fp(&o->x, x3);
        bam_st_fp(&o->y, y3);
    }
}

// final: ONE THREAD per MSM adds its `per` (<= 96) remaining points one after the other.  There are 128 MSMs per
// blob, so the whole batch is resident at once and the kernel lasts one chain of `per` mixed additions; a warp
// per MSM with a shared-memory tree (the shape of fk20_msm_kernel's tail) took 3.5 ms for 256 blobs at
// per = 88 -- few resident CTAs, each waiting on a five-level tree -- this takes 0.9 ms.
__global__ void __launch_bounds__(BAM_THREADS) bam_fk_final_kernel(G1* __restrict__ u_brp, const G1Affine* __restrict__ in, uint32_t per, uint64_t total_msm) {
    const uint64_t msm = (uint64_t)blockIdx.x * BAM_THREADS + threadIdx.x;
    if (msm >= total_msm) return;
    const G1Affine* src = in + msm * (size_t)per;
    G1 acc = g1_inf();
#pragma unroll 1
    for (uint32_t k = 0; k < per; k++) {
        G1Affine a;
        a.x = bam_ld_fp(&src[k].x);
        a.y = bam_ld_fp(&src[k].y);
        g1_madd_to(acc, a, false);
    }
    const uint64_t blob = msm >> 7;
    const int slot = (int)(__brev((uint32_t)(msm & 127)) >> 25);
    g1_store(u_brp + blob * 128 + slot, acc);
}

size_t fk20_msm_affine_workspace_bytes(uint64_t n, int c) {
    const FkGeom g = fk_geom(c);
    const size_t per = (size_t)32 * g.w;
    return n * 128 * (per + per / 2) * sizeof(G1Affine);
}

// u_brp as launch_fk20_msm (fk20.cu); the table (Ctx::fk_table) must exist
int launch_fk20_msm_affine(Launch& L, G1* u_brp, const uint32_t* S, uint64_t n, void* workspace) {
    if (n == 0) return RET_OK;
    Ctx* c = L.ctx;
    const FkGeom g = fk_geom(c->fk_c);
    const uint64_t total_msm = n * 128;
    uint32_t per = (uint32_t)(32 * g.w);
    G1Affine* a = (G1Affine*)workspace;
    G1Affine* b = a + total_msm * (size_t)per;
    bam_fk_level1_kernel<<<(unsigned)((total_msm * 16 + BAM_THREADS - 1) / BAM_THREADS), BAM_THREADS, 0, L.stream>>>(a, S, (const G1Affine*)c->fk_table, total_msm, g);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "fk20_msm_level1");
    int levels = 0;
    while (per > 96 && (per & 1u) == 0) {
        const uint64_t half = total_msm * (per / 2);
        if (half >= (1ull << 32)) return RET_ERROR;
        const uint64_t threads = (half + BAM_SEG - 1) / BAM_SEG;
        bam_level_kernel<<<dim3((unsigned)((threads + BAM_THREADS - 1) / BAM_THREADS), 1), BAM_THREADS, 0, L.stream>>>(b, a, (uint32_t)half);
        KZG_CUDA_TRY(cudaGetLastError());
        G1Affine* tmp = a;
        a = b;
        b = tmp;
        per /= 2;
        levels++;
    }
    L.count(levels, "fk20_msm_levels");
    bam_fk_final_kernel<<<(unsigned)((total_msm + BAM_THREADS - 1) / BAM_THREADS), BAM_THREADS, 0, L.stream>>>(u_brp, a, per, total_msm);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "fk20_msm_final");
    return RET_OK;
}

size_t msm_affine_workspace_bytes(uint64_t n, int c) {
    const FkGeom g = fk_geom(c);
    const size_t n1 = (size_t)2048 * g.w;
    return n * (n1 + n1 / 2) * sizeof(G1Affine);
}

int launch_msm_affine(Launch& L, G1* result, const uint8_t* scalars, bool big_endian_bytes, uint64_t n, int* d_bad, void* workspace) {
    if (n == 0) return RET_OK;
    Ctx* c = L.ctx;
    const FkGeom g = fk_geom(c->commit_c);
    uint32_t count = (uint32_t)(2048 * g.w);
    G1Affine* a = (G1Affine*)workspace;
    G1Affine* b = a + n * (size_t)count;
    bam_level1_kernel<<<dim3(1024 / BAM_THREADS, (unsigned)n), BAM_THREADS, 0, L.stream>>>(a, scalars, big_endian_bytes, (const G1Affine*)c->commit_table, d_bad, g);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "msm_affine_level1");
    int levels = 0;
    while (count > (uint32_t)BAM_STOP && (count & 1u) == 0) {
        const uint32_t half = count / 2;
        const unsigned threads = (half + BAM_SEG - 1) / BAM_SEG;
        bam_level_kernel<<<dim3((threads + BAM_THREADS - 1) / BAM_THREADS, (unsigned)n), BAM_THREADS, 0, L.stream>>>(b, a, half);
        KZG_CUDA_TRY(cudaGetLastError());
        G1Affine* tmp = a;
        a = b;
        b = tmp;
        count = half;
        levels++;
    }
    L.count(levels, "msm_affine_levels");
    bam_final_kernel<<<(unsigned)n, BAM_THREADS, 0, L.stream>>>(result, a, count);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "msm_affine_final");
    return RET_OK;
}

}  // namespace kzg
