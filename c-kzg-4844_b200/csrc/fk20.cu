// FK20 cell proofs, G1 side: fixed-base tables, the 128 x MSM(64) and the two G1 FFTs.
//
// Replaces (paths relative to the reference tree):
//   init_fk20_multi_settings .............. src/setup/setup.c:238-330 (x_ext_fft_columns, tables)
//   compute_fk20_cell_proofs, phases 1b-2 .. src/eip7594/fk20.c:213-270 (MSMs, g1_ifft_unscaled, g1_fft)
//   g1_fft_fast ........................... src/eip7594/fft.c:164-185
//   blst_p1s_mult_wbits(_precompute) ...... blst/src/multi_scalar.c:133-262
//
// B200 design: the 8192 bases X^[j][i] are fixed, so setup stores every small multiple of every
// window shift, T[p][w][m] = (m+1) 2^(cw) X^_p (c = 12: 35 GB of the 180 GB HBM).  An MSM(64) is then a
// pure gather-and-add of 64 x ceil(256/c) table points: one warp per MSM, two base points per lane, a
// shared-memory tree at the end -- no buckets, no doublings, no sorting.  The G1 FFTs live in
// fk20_fft.cu; ordering is arranged so that no permutation pass exists (MSM j stores to slot brp7(j);
// inverse DIT gives natural order; forward DIF leaves the proofs in the bit-reversed order the API
// returns).
#include <stdlib.h>

#include "call.h"
#include "cells.h"
#include "g1_hot.cuh"

namespace kzg {

__device__ __forceinline__ G1 ld_g1(const G1* p) {
    G1 a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) d[i] = q[i];
    return a;
}
__device__ __forceinline__ void st_g1(G1* p, const G1& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* d = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) q[i] = d[i];
}
__device__ __forceinline__ G1Affine ld_affine(const G1Affine* p) {
    G1Affine a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = __ldg(q + i);
    return a;
}
__device__ __forceinline__ Fr ld_fr2(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ int brp7(int v) { return (int)(__brev((uint32_t)v) >> 25); }

// ------------------------------------------------------------------------------------------------
// setup: X^ columns and the window tables
// ------------------------------------------------------------------------------------------------
// xin[offset][k] = g1_monomial[4096 - 64 - 1 - offset - 64 k] for k < 63, infinity otherwise (setup.c:272-282)
__global__ void fk_gather_x_kernel(G1* __restrict__ xin, const G1Affine* __restrict__ g1_monomial) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 64 * 128) return;
    int offset = e >> 7, k = e & 127;
    G1 p = g1_inf();
    if (k < 63) p = g1_from_affine(g1_monomial[N_BLOB - 64 - 1 - offset - 64 * k]);
    st_g1(xin + e, p);
}
// FFT output slot q of vector `offset` is row brp7(q):  xhat[row * 64 + offset] (affine)
__global__ void fk_xhat_affine_kernel(G1Affine* __restrict__ xhat, const G1* __restrict__ fft_out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 64 * 128) return;
    int offset = e >> 7, q = e & 127;
    G1 p = ld_g1(fft_out + e);
    xhat[brp7(q) * 64 + offset] = g1_to_affine(p);
}
// bases[p][w] = 2^(cw) * xhat[p], affine
__global__ void fk_bases_kernel(G1Affine* __restrict__ bases, const G1Affine* __restrict__ xhat, int npts, const FkGeom g) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    G1Affine a = xhat[p];
    bases[(size_t)p * g.w] = a;
    G1 acc = g1_from_affine(a);
#pragma unroll 1
    for (int w = 1; w < g.w; w++) {
#pragma unroll 1
        for (int k = 0; k < g.c; k++) g1_dbl_to(acc);
        G1Affine q = g1_to_affine(acc);
        bases[(size_t)p * g.w + w] = q;
        acc = g1_from_affine(q);
    }
}
// table[(p*W + w)*M + m] = (m+1) * bases[p][w], affine; one thread per (p, w), batches of 16
// converted with one inversion each (Montgomery's trick)
constexpr int FK_BATCH = 16;
__global__ void __launch_bounds__(64) fk_multiples_kernel(G1Affine* __restrict__ table, const G1Affine* __restrict__ bases, int npts, const FkGeom g) {
    size_t pw = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pw >= (size_t)npts * g.w) return;
    const G1Affine base = bases[pw];
    G1Affine* dst = table + pw * g.m;
    const int FK_M = g.m;
    G1 run = g1_from_affine(base);
    G1 pts[FK_BATCH];
    Fp pre[FK_BATCH];
#pragma unroll 1
    for (int m0 = 0; m0 < FK_M; m0 += FK_BATCH) {
        Fp prod = Fp::one();
#pragma unroll 1
        for (int k = 0; k < FK_BATCH; k++) {
            pts[k] = run;
            pre[k] = prod;
            Fp zc = mul(run.zz, run.zzz);
            if (is_zero(zc)) zc = Fp::one();  // infinity: keep the chain invertible, emit (0,0) below
            prod = mul(prod, zc);
            g1_madd_to(run, base, false);
        }
        Fp inv = fp_inv(prod);
#pragma unroll 1
        for (int k = FK_BATCH - 1; k >= 0; k--) {
            G1Affine a = g1a_inf();
            const G1& pt = pts[k];
            Fp zc = mul(pt.zz, pt.zzz);
            if (!is_zero(zc)) {
                Fp iz = mul(inv, pre[k]);  // 1 / (zz * zzz)
                inv = mul(inv, zc);
                a.x = mul(pt.x, mul(iz, pt.zzz));
                a.y = mul(pt.y, mul(iz, pt.zz));
            }
            dst[m0 + k] = a;
        }
    }
}

// table[(p*W + w)*M + m] = (m+1) 2^(cw) pts[p] for npts affine points (fixed-base tables of FK20 and of the
// direct commitment MSM, msm_direct.cu)
int launch_fixed_base_table(Launch& L, G1Affine* table, const G1Affine* pts, int npts, const FkGeom& g) {
    G1Affine* bases = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&bases, (size_t)npts * g.w * sizeof(G1Affine), L.stream));
    fk_bases_kernel<<<(npts + 63) / 64, 64, 0, L.stream>>>(bases, pts, npts, g);
    KZG_CUDA_TRY(cudaGetLastError());
    fk_multiples_kernel<<<(unsigned)(((size_t)npts * g.w + 63) / 64), 64, 0, L.stream>>>(table, bases, npts, g);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaFreeAsync(bases, L.stream));
    L.count(2, "fixed_base_table");
    return RET_OK;
}

// `max_c`: widest window to try; a failed allocation falls back to the next smaller table
static int fk20_build(Launch& L, Ctx* c, int max_c) {
    static const int widths[3] = {12, 10, 8};
    c->fk_table = nullptr;
    for (int k = 0; k < 3; k++) {
        if (widths[k] > max_c) continue;
        const FkGeom g = fk_geom(widths[k]);
        if (cudaMalloc((void**)&c->fk_table, g.table_points() * sizeof(G1Affine)) == cudaSuccess) {
            c->fk_c = widths[k];
            break;
        }
        (void)cudaGetLastError();
        c->fk_table = nullptr;
    }
    if (!c->fk_table) return RET_MALLOC;
    const FkGeom g = fk_geom(c->fk_c);
    G1* xin = nullptr;
    G1Affine* xhat = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&xin, 64 * 128 * sizeof(G1), L.stream));
    KZG_CUDA_TRY(cudaMallocAsync((void**)&xhat, FK_POINTS * sizeof(G1Affine), L.stream));
    fk_gather_x_kernel<<<64 * 128 / 128, 128, 0, L.stream>>>(xin, c->g1_monomial);
    KZG_CUDA_TRY(cudaGetLastError());
    {
        int rc = g1_fft128_run(L, xin, 64, false);
        if (rc) return rc;
    }
    fk_xhat_affine_kernel<<<64 * 128 / 64, 64, 0, L.stream>>>(xhat, xin);
    KZG_CUDA_TRY(cudaGetLastError());
    {
        int rc = launch_fixed_base_table(L, (G1Affine*)c->fk_table, xhat, FK_POINTS, g);
        if (rc) return rc;
    }
    KZG_CUDA_TRY(cudaFreeAsync(xin, L.stream));
    KZG_CUDA_TRY(cudaFreeAsync(xhat, L.stream));
    L.count(2, "fk20_setup");
    KZG_CUDA_TRY(cudaStreamSynchronize(L.stream));
    return RET_OK;
}

// X^ columns (init_fk20_multi_settings, setup.c:238-330) + the window tables, on first use of an API that
// computes cell proofs
// A failed build (not only a failed allocation) releases the table and retries with the next smaller window;
// if even the 8-bit table cannot be built the error is reported but NOT remembered: the next call tries again.
int fk20_ensure(Ctx* c) {
    if (c->fk_ready.load(std::memory_order_acquire)) return RET_OK;
    std::lock_guard<std::mutex> g(c->fk_mu);
    if (c->fk_ready.load(std::memory_order_relaxed)) return RET_OK;
    int rc = RET_ERROR;
    for (int max_c = plan_fk_window(); max_c >= 8; max_c = c->fk_c - 2) {
        {
            Call call(c);
            if (!call.ok) return RET_ERROR;
            Launch L = call.launch();
            rc = fk20_build(L, c, max_c);
        }
        if (rc == RET_OK) {
            c->fk_ready.store(true, std::memory_order_release);
            return RET_OK;
        }
        (void)cudaGetLastError();
        if (!c->fk_table) break;  // nothing could be allocated at all
        cudaFree(c->fk_table);
        c->fk_table = nullptr;
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// 128 x MSM(64) per blob: one warp per MSM
// ------------------------------------------------------------------------------------------------
constexpr int FM_WARPS = 4;
constexpr int FM_AHEAD = 4;  // table entries requested ahead of the addition that uses them

__global__ void __launch_bounds__(32 * FM_WARPS, 3) fk20_msm_kernel(G1* __restrict__ u_brp, const uint32_t* __restrict__ S, const G1Affine* __restrict__ table, uint64_t total,
                                                                 const FkGeom g) {
    __shared__ G1 sh[FM_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t msm = (uint64_t)blockIdx.x * FM_WARPS + warp;  // = blob * 128 + j
    const bool active = msm < total;
    G1 acc = g1_inf();
    if (active) {
        const int j = (int)(msm & 127);
        // digits of the lane's two scalars first (signed c-bit, the carry runs up the windows), so that the
        // table entry of step t + FM_AHEAD can be pulled into L2 while step t is being added: the gathers are
        // random 96-byte reads over a table of up to 35 GB
        uint16_t dig[64];  // magnitude | sign << 15, index = half * W + window
        const G1Affine* tp[2];
        const uint32_t dmask = (1u << g.c) - 1u, dfull = 1u << g.c;
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int i = lane + 32 * h;
            const uint4* sp = reinterpret_cast<const uint4*>(S + (msm * 64 + i) * 8);
            uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
            uint32_t s[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0u};
            tp[h] = table + ((size_t)(j * 64 + i) * g.w) * g.m;
            uint32_t carry = 0;
#pragma unroll 1
            for (int w = 0; w < g.w; w++) {
                const int o = w * g.c;  // the scalar is below 2^255: the top digit absorbs the carry
                uint32_t d = (__funnelshift_r(s[o >> 5], s[(o >> 5) + 1], o & 31) & dmask) + carry;
                bool negd = d > (uint32_t)g.m;
                carry = negd ? 1u : 0u;
                uint32_t mag = negd ? (dfull - d) : d;
                dig[h * g.w + w] = (uint16_t)(mag | (negd ? 0x8000u : 0u));
            }
        }
        const int steps = 2 * g.w;
        auto entry = [&](int t) -> const G1Affine* {
            const int h = t >= g.w ? 1 : 0;
            return tp[h] + (size_t)(t - h * g.w) * g.m + ((dig[t] & 0x7fffu) - 1u);
        };
        auto pull = [&](int t) {
            if (t < steps && (dig[t] & 0x7fffu) != 0) {
                const char* e = reinterpret_cast<const char*>(entry(t));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(e));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(e + 64));
            }
        };
#pragma unroll 1
        for (int t = 0; t < FM_AHEAD; t++) pull(t);
#pragma unroll 1
        for (int t = 0; t < steps; t++) {
            pull(t + FM_AHEAD);
            if ((dig[t] & 0x7fffu) != 0) {
                G1Affine a = ld_affine(entry(t));
                g1_madd_nl(acc, a, (dig[t] & 0x8000u) != 0);
            }
        }
    }
    sh[warp][lane] = acc;
    __syncwarp();
#pragma unroll 1
    for (int s = 16; s > 0; s >>= 1) {
        if (lane < s) {
            G1 x = sh[warp][lane], y = sh[warp][lane + s];
            g1_add_to(x, y);
            sh[warp][lane] = x;
        }
        __syncwarp();
    }
    if (active && lane == 0) {
        const uint64_t blob = msm >> 7;
        st_g1(u_brp + blob * 128 + brp7((int)(msm & 127)), sh[warp][0]);
    }
}

int launch_fk20_msm(Launch& L, G1* u_brp, const uint32_t* S, uint64_t n) {
    if (!n) return RET_OK;
    {
        int rc = fk20_ensure(L.ctx);
        if (rc) return rc;
    }
    static const int affine_min = getenv("CKZG_B200_FK_AFFINE_MIN") ? atoi(getenv("CKZG_B200_FK_AFFINE_MIN")) : 8;  // 0 = never
    if (affine_min > 0 && n >= (uint64_t)affine_min) {
        // 13 MB of level scratch per blob: slices of 512 blobs (6.6 GB) fill the GPU many times over, larger batches
        // (the FFTs behind take up to 2048 at once) reuse the same scratch slice by slice
        const uint64_t SL = 512;
        const uint64_t slice = n < SL ? n : SL;
        void* ws = nullptr;
        KZG_CUDA_TRY(cudaMallocAsync(&ws, fk20_msm_affine_workspace_bytes(slice, L.ctx->fk_c), L.stream));
        int rc = RET_OK;
        for (uint64_t off = 0; off < n && rc == RET_OK; off += slice) {
            const uint64_t m = (n - off < slice) ? n - off : slice;
            rc = launch_fk20_msm_affine(L, u_brp + off * 128, S + off * 128 * 64 * 8, m, ws);
        }
        cudaFreeAsync(ws, L.stream);
        return rc;
    }
    uint64_t total = n * 128;
    fk20_msm_kernel<<<(unsigned)((total + FM_WARPS - 1) / FM_WARPS), 32 * FM_WARPS, 0, L.stream>>>(u_brp, S, (const G1Affine*)L.ctx->fk_table, total, fk_geom(L.ctx->fk_c));
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "fk20_msm");
    return RET_OK;
}

int launch_fk20_g1_ffts(Launch& L, G1* proofs, G1* u_brp, uint64_t n) {
    if (!n) return RET_OK;
    int rc = g1_fft128_run(L, u_brp, n, true);  // in place
    if (rc) return rc;
    KZG_CUDA_TRY(cudaMemcpyAsync(proofs, u_brp, n * 128 * sizeof(G1), cudaMemcpyDeviceToDevice, L.stream));
    L.count(14, "fk20_g1_ffts");
    return RET_OK;
}

}  // namespace kzg
