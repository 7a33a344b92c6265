// MSM over the 4096 Lagrange setup points WITHOUT buckets: every small multiple of every window shift
// of every base is stored, D[i][w][m] = (m+1) 2^(cw) L_i (affine), so a commitment is a pure
// gather-and-add of 4096 x ceil(256/c) table points.
//
// Replaces (paths relative to the reference tree) g1_lincomb_fast (src/common/lincomb.c:65 ->
// blst_p1s_mult_pippenger, blst/src/multi_scalar.c:370) for the fixed bases of blob_to_kzg_commitment
// (src/eip4844/eip4844.c:264) and of the quotient commitment of compute_kzg_proof_impl (:417-494).
//
// Against the bucket form (msm.cu: 24 windows x 4096 additions, a counting sort before and a 1024-bucket
// reduction after): c = 14 needs 19 x 4096 additions and no sort / reduction at all, for 61 GB of the
// 180 GB HBM; c = 13: 20 windows, 32 GB; c = 12: 22 windows, 18 GB.  The width is chosen per context from
// the device memory that is free when the table is first needed (api.cu plan_commit_window); without room the
// bucket form stays in use.
//
// Layout: 8 CTAs of 128 threads per blob (32 when fewer than 64 blobs must fill the GPU); a thread owns 4
// (1) scalars (coalesced 32-byte loads), walks their windows with one XYZZ accumulator (table entries pulled
// into L2 a few steps ahead), the CTA folds its 128 accumulators through shared memory, and a second small
// kernel adds the partial sums of a blob (one warp per blob).
#include "call.h"
#include "cells.h"
#include "g1_hot.cuh"

namespace kzg {

constexpr int MD_THREADS = 128;
constexpr int MD_MAX_CTAS = 32;  // per blob: 4 scalars per thread (8 CTAs) for batches, 1 (32 CTAs) when few blobs must fill the GPU
constexpr int MD_AHEAD = 4;

__device__ __forceinline__ uint32_t md_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ G1Affine md_ld_affine(const G1Affine* p) {
    G1Affine a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = __ldg(q + i);
    return a;
}

__global__ void __launch_bounds__(MD_THREADS, 3) msm_direct_kernel(G1* __restrict__ partial, const uint8_t* __restrict__ scalars, bool big_endian, const G1Affine* __restrict__ table,
                                                                   int* __restrict__ bad, const FkGeom g, int per_thread) {
    __shared__ G1 sh[MD_THREADS];
    const int t = threadIdx.x;
    const uint64_t blob = blockIdx.y;
    const uint32_t dmask = (1u << g.c) - 1u, dfull = 1u << g.c;
    G1 acc = g1_inf();
    uint16_t dig[32];  // magnitude | sign << 15 per window (W <= 32, magnitude <= 2^13)
#pragma unroll 1
    for (int k = 0; k < per_thread; k++) {
        const int i = (blockIdx.x * per_thread + k) * MD_THREADS + t;  // element index inside the blob
        const uint4* sp = reinterpret_cast<const uint4*>(scalars + (blob * N_BLOB + i) * 32);
        const uint4 a = __ldg(sp), b = __ldg(sp + 1);
        uint32_t s[9];
        if (big_endian) {  // wire form: 32 bytes big-endian, must be canonical (bytes_to_bls_field, src/common/bytes.c:64)
            s[0] = md_bswap(b.w); s[1] = md_bswap(b.z); s[2] = md_bswap(b.y); s[3] = md_bswap(b.x);
            s[4] = md_bswap(a.w); s[5] = md_bswap(a.z); s[6] = md_bswap(a.y); s[7] = md_bswap(a.x);
            if (limbs_geq<8>(s, FR_MOD)) {
                if (bad) bad[blob] = 1;
                continue;
            }
        } else {
            s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
            s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
        }
        s[8] = 0;
        uint32_t carry = 0;
#pragma unroll 1
        for (int w = 0; w < g.w; w++) {
            const int o = w * g.c;  // the scalar is below 2^255: the top digit absorbs the carry
            uint32_t d = (__funnelshift_r(s[o >> 5], s[(o >> 5) + 1], o & 31) & dmask) + carry;
            const bool negd = d > (uint32_t)g.m;
            carry = negd ? 1u : 0u;
            const uint32_t mag = negd ? (dfull - d) : d;
            dig[w] = (uint16_t)(mag | (negd ? 0x8000u : 0u));
        }
        const G1Affine* tp = table + ((size_t)i * g.w) * g.m;
        auto pull = [&](int w) {
            if (w < g.w && (dig[w] & 0x7fffu) != 0) {
                const char* e = reinterpret_cast<const char*>(tp + (size_t)w * g.m + ((dig[w] & 0x7fffu) - 1u));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(e));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(e + 64));
            }
        };
#pragma unroll 1
        for (int w = 0; w < MD_AHEAD; w++) pull(w);
#pragma unroll 1
        for (int w = 0; w < g.w; w++) {
            pull(w + MD_AHEAD);
            const uint32_t mag = dig[w] & 0x7fffu;
            if (mag != 0) {
                G1Affine e = md_ld_affine(tp + (size_t)w * g.m + (mag - 1u));
                g1_madd_nl(acc, e, (dig[w] & 0x8000u) != 0);
            }
        }
    }
    sh[t] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = MD_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) {
            G1 x = sh[t], y = sh[t + s];
            g1_add_to(x, y);
            sh[t] = x;
        }
        __syncthreads();
    }
    if (t < 12) reinterpret_cast<uint4*>(partial + blob * gridDim.x + blockIdx.x)[t] = reinterpret_cast<const uint4*>(&sh[0])[t];
}

// result[blob] = sum of the blob's `ctas` partial sums: one warp per blob, a tree through shared memory
__global__ void __launch_bounds__(32) msm_direct_fold_kernel(G1* __restrict__ result, const G1* __restrict__ partial, int ctas) {
    __shared__ G1 sh[MD_MAX_CTAS];
    const int lane = threadIdx.x;
    const uint64_t blob = blockIdx.x;
    if (lane < ctas) sh[lane] = partial[blob * ctas + lane];
    __syncwarp();
#pragma unroll 1
    for (int s = ctas / 2; s > 0; s >>= 1) {
        if (lane < s) {
            G1 x = sh[lane], y = sh[lane + s];
            g1_add_to(x, y);
            sh[lane] = x;
        }
        __syncwarp();
    }
    if (lane < 12) reinterpret_cast<uint4*>(result + blob)[lane] = reinterpret_cast<const uint4*>(&sh[0])[lane];
}

size_t msm_direct_workspace_bytes(uint64_t n) { return n * MD_MAX_CTAS * sizeof(G1); }

int launch_msm_direct(Launch& L, G1* result, const uint8_t* scalars, bool big_endian_bytes, uint64_t n, int* d_bad, void* workspace) {
    if (n == 0) return RET_OK;
    Ctx* c = L.ctx;
    G1* partial = (G1*)workspace;
    const int per_thread = n >= 64 ? 4 : 1;
    const int ctas = N_BLOB / (MD_THREADS * per_thread);
    msm_direct_kernel<<<dim3(ctas, (unsigned)n), MD_THREADS, 0, L.stream>>>(partial, scalars, big_endian_bytes, (const G1Affine*)c->commit_table, d_bad, fk_geom(c->commit_c),
                                                                            per_thread);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "msm_direct");
    msm_direct_fold_kernel<<<(unsigned)n, 32, 0, L.stream>>>(result, partial, ctas);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "msm_direct_fold");
    return RET_OK;
}

// Never fails the caller: when the table cannot be allocated OR its build fails (a transient error of a
// co-tenant, say) the table is released and the bucket form (msm.cu, msm_table) serves this context, as
// include/ckzg_b200.h promises.
int msm_direct_ensure(Ctx* c) {
    std::call_once(c->commit_once, [c] {
        c->commit_c = plan_commit_window();
        if (c->commit_c == 0) return;  // no room: the bucket form (msm.cu) serves
        const FkGeom g = fk_geom(c->commit_c);
        if (cudaMalloc((void**)&c->commit_table, g.points_for(N_BLOB) * sizeof(G1Affine)) != cudaSuccess) {
            (void)cudaGetLastError();
            c->commit_table = nullptr;
            c->commit_c = 0;
            return;
        }
        int rc = RET_ERROR;
        {
            Call call(c);
            if (call.ok) {
                Launch L = call.launch();
                rc = launch_fixed_base_table(L, (G1Affine*)c->commit_table, c->g1_lagrange_brp, N_BLOB, g);
                if (rc == RET_OK && cudaStreamSynchronize(call.stream) != cudaSuccess) rc = RET_ERROR;
            }
        }
        if (rc != RET_OK) {
            (void)cudaGetLastError();
            cudaFree(c->commit_table);
            c->commit_table = nullptr;
            c->commit_c = 0;
        }
    });
    return RET_OK;
}

}  // namespace kzg
