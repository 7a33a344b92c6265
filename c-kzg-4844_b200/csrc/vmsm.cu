// Variable-base bucket MSM for the random linear combination of verify_blob_kzg_proof_batch.
//
// Replaces the three g1_lincomb_naive calls + g1_mul of verify_kzg_proof_batch
// (src/eip4844/eip4844.c:724-747; g1_lincomb_naive src/common/lincomb.c:34-50) for batches:
//     A = sum r^i proof_i,   B = sum (r^i z_i) proof_i + sum r^i C_i - [sum r^i y_i] G1.
// Same group elements, hence the same pairing verdict.
//
// Why not one scalar multiplication per point (rlc_points_kernel): a 255-bit multiplication is a chain
// of ~2000 dependent Fp products on one lane (2.5 ms) and it can only start once the challenge r is
// known, i.e. it sits in the serial tail of every call.  Here the part of that chain that does not
// depend on r -- the doublings -- is taken from work that happens BEFORE r anyway:
//   * the subgroup test of every commitment / proof multiplies by |z| twice; done LSB-first
//     (g1.cuh g1a_validate_levels) it walks the doubling chains of P and of Q = [|z|]P and leaves
//     table[j][i] = 2^(8j) P_i (j < 9) and 2^(8(j-9)) Q_i (j = 9..17) behind, for free;
//   * once r is known: scalars r^i, r^i z_i (rlc_vmsm_scalars_kernel) are written in base |z|,
//     k = a0 + a1|z| + a2|z|^2 + a3|z|^3 with 64-bit digits, so that
//     [k]P = a0 P + a1 Q + a2 (-phi(P)) + a3 (-phi(Q))   (phi(x,y) = (beta x, y) acts as -z^2),
//     and each 64-bit digit is recoded into signed bytes; every non-zero byte d of (point i, level j,
//     quarter q) is one table entry that belongs in bucket |d| -- all levels share ONE set of 128
//     buckets because the shifts are already applied;
//   * vmsm_hist_kernel / vmsm_scatter_kernel: counting sort by bucket over 64 CTAs per MSM;
//   * vmsm_accumulate_kernel: one thread per <= 8-entry slice of a bucket list (XYZZ + XYZZ adds);
//   * vmsm_combine_kernel: one CTA per bucket folds its slices; vmsm_reduce_kernel: sum_b (b+1) B_b by
//     a suffix scan and a tree over the 128 buckets.
// The serial depth after r is ~35 group additions instead of ~190 doublings/additions, and no
// doubling at all.
#define KZG_FP_MUL_OUTLINE 1
#include "g1_glv.cuh"
#include "g1_quad.cuh"
#include "vmsm.cuh"

namespace kzg {

__device__ __forceinline__ G1 vload_g1(const G1* p) {
    G1 a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) d[i] = q[i];
    return a;
}
__device__ __forceinline__ void vstore_g1(G1* p, const G1& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* d = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) q[i] = d[i];
}

// ------------------------------------------------------------------------------------------------
// after r: scalars -> base-|z| digits
// ------------------------------------------------------------------------------------------------
// hB layout (index = 4 * point + quarter, 8 bytes each): points 0..n-1 proofs with r^i z_i, points
// n..2n-1 commitments with r^i, point 2n = -G with sum r^i y_i.  MSM A reads the r^i segment with point
// base 0.
// `first`: this call's points are tuples [first, first + n) of a larger batch (sharded verification): weights r^(first + i)
__global__ void rlc_vmsm_scalars_kernel(uint32_t* __restrict__ hB, Fr* __restrict__ ty, const Fr* __restrict__ z, const Fr* __restrict__ y, const Digest8 digest,
                                        uint32_t n, uint64_t first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr p = Fr::one();
    {
        Fr base = fr_from_digest_words(digest.h);
        uint64_t e = first + i;
        while (e) {
            if (e & 1) p = mul(p, base);
            base = sqr(base);
            e >>= 1;
        }
    }
    uint32_t s[8];
    from_mont<FrTag>(s, p);
    store_halves(hB, (size_t)(n + i), s);
    from_mont<FrTag>(s, mul(p, z[i]));
    store_halves(hB, (size_t)i, s);
    ty[i] = mul(p, y[i]);
}

__global__ void __launch_bounds__(256) rlc_vmsm_ysum_kernel(uint32_t* __restrict__ hB, const Fr* __restrict__ ty, uint32_t n) {
    __shared__ Fr sh[256];
    const int t = threadIdx.x;
    Fr s = Fr::zero();
    for (uint32_t i = t; i < n; i += 256) s = add(s, ty[i]);
    sh[t] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (t < k) sh[t] = add(sh[t], sh[t + k]);
        __syncthreads();
    }
    if (t == 0) {
        uint32_t v[8];
        from_mont<FrTag>(v, sh[0]);
        store_halves(hB, (size_t)(2 * n), v);
    }
}

// ------------------------------------------------------------------------------------------------
// counting sort of the digits by bucket: one CTA per MSM
// ------------------------------------------------------------------------------------------------
// calls f(level j, bucket b, negative) for every non-zero signed byte of the balanced digit v
// (magnitude below 0x6a00.. in bits 0..62, sign in bit 63: the top byte never carries out)
template <class Fn>
__device__ __forceinline__ void for_each_byte_digit(const uint2 v, Fn f) {
    const uint32_t w[2] = {v.x, v.y & 0x7fffffffu};
    const bool sgn = (v.y >> 31) != 0;
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t d = ((w[j >> 2] >> ((j & 3) * 8)) & 0xffu) + carry;
        const bool negd = d > (uint32_t)VNB;
        carry = negd ? 1u : 0u;
        const uint32_t mag = negd ? (256u - d) : d;  // 1..128 (or 0)
        if (mag != 0) f(j, mag - 1u, negd != sgn);
    }
}

// Two launches over VSORT_CTAS CTAs per MSM, each CTA owning a contiguous slice of the half-scalars:
//   vmsm_hist_kernel   -- per-CTA bucket histogram -> ctahist[cta][bucket]
//   vmsm_scatter_kernel -- every CTA derives the bucket starts (sum over CTAs, scan over buckets), its own
//                          offset inside each bucket (CTAs before it) and the offsets of its warps (a second
//                          count in shared memory), then writes its entries; the work-item tables are filled
//                          cooperatively (CTA c takes the buckets b = c mod VSORT_CTAS).
// (The first version sorted each MSM in ONE 1024-thread CTA: 179 us at n = 4096, almost all of it
// shared-memory atomic latency -- ncu r01t: short_scoreboard 11.5 stalls per issue.)
__device__ __forceinline__ void vsort_slice(uint32_t& h0, uint32_t& h1, uint32_t nh) {
    const uint32_t per = (nh + VSORT_CTAS - 1) / VSORT_CTAS;
    h0 = blockIdx.x * per;
    h1 = h0 + per < nh ? h0 + per : nh;
    if (h0 > nh) h0 = nh;
}

__global__ void __launch_bounds__(VSORT_THREADS) vmsm_hist_kernel(const __grid_constant__ VmsmJobs jobs) {
    __shared__ uint32_t cnt[VNB];
    const VmsmJob& J = jobs.j[blockIdx.y];
    const int tid = threadIdx.x;
    if (tid < VNB) cnt[tid] = 0;
    __syncthreads();
    uint32_t h0, h1;
    vsort_slice(h0, h1, J.nh);
    const uint2* hv = reinterpret_cast<const uint2*>(J.halves);
    for (uint32_t h = h0 + tid; h < h1; h += VSORT_THREADS) {
        for_each_byte_digit(hv[h], [&](int, uint32_t b, bool) { atomicAdd(&cnt[b], 1u); });
    }
    __syncthreads();
    if (tid < VNB) J.ctahist[blockIdx.x * VNB + tid] = cnt[tid];
}

__global__ void __launch_bounds__(VSORT_THREADS) vmsm_scatter_kernel(const __grid_constant__ VmsmJobs jobs, uint32_t npts) {
    __shared__ uint32_t cnt[VSORT_WARPS][VNB];
    __shared__ uint32_t bstart[VNB + 1];  // global start of every bucket
    __shared__ uint32_t mine[VNB];        // where this CTA's entries start inside the bucket
    __shared__ uint32_t istart[VNB + 1];
    const VmsmJob& J = jobs.j[blockIdx.y];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < VSORT_WARPS * VNB; i += VSORT_THREADS) (&cnt[0][0])[i] = 0;
    if (tid < VNB) {
        uint32_t total = 0, before = 0;
        for (int c = 0; c < VSORT_CTAS; c++) {
            const uint32_t v = J.ctahist[c * VNB + tid];
            if (c < (int)blockIdx.x) before += v;
            total += v;
        }
        bstart[tid] = total;
        mine[tid] = before;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0, irun = 0;
        for (int b = 0; b < VNB; b++) {
            const uint32_t c = bstart[b];
            bstart[b] = run;
            istart[b] = irun;
            run += c;
            irun += c ? (c + VCAP - 1) / VCAP : 1u;
        }
        bstart[VNB] = run;
        istart[VNB] = irun;
    }
    uint32_t h0, h1;
    vsort_slice(h0, h1, J.nh);
    const uint2* hv = reinterpret_cast<const uint2*>(J.halves);
    for (uint32_t h = h0 + tid; h < h1; h += VSORT_THREADS) {
        for_each_byte_digit(hv[h], [&](int, uint32_t b, bool) { atomicAdd(&cnt[warp][b], 1u); });
    }
    __syncthreads();
    if (tid < VNB) {  // exclusive prefix over the warps of this CTA, on top of the CTA's own offset
        uint32_t run = bstart[tid] + mine[tid];
        for (int w = 0; w < VSORT_WARPS; w++) {
            const uint32_t c = cnt[w][tid];
            cnt[w][tid] = run;
            run += c;
        }
    }
    if (blockIdx.x == 0 && tid <= VNB) {
        J.starts[tid] = bstart[tid];
        J.item_start[tid] = istart[tid];
    }
    for (int b = blockIdx.x; b < VNB; b += VSORT_CTAS)
        for (uint32_t k = istart[b] + tid; k < istart[b + 1]; k += VSORT_THREADS) J.item_bucket[k] = (uint32_t)b;
    __syncthreads();
    for (uint32_t h = h0 + tid; h < h1; h += VSORT_THREADS) {
        const uint32_t pt = h >> 2, qtr = h & 3u, phi = qtr >> 1, lvl0 = (qtr & 1u) * VW;
        for_each_byte_digit(hv[h], [&](int j, uint32_t b, bool negd) {
            const uint32_t pos = atomicAdd(&cnt[warp][b], 1u);
            // quarters 2, 3 use the bases -phi(P), -phi(Q): (beta X, -Y)
            J.entries[pos] = ((lvl0 + (uint32_t)j) * npts + pt) | (phi ? 0x40000000u : 0u) | ((negd != (phi != 0)) ? 0x80000000u : 0u);
        });
    }
}

// ------------------------------------------------------------------------------------------------
// bucket sums
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VACC_THREADS) vmsm_accumulate_kernel(const __grid_constant__ VmsmJobs jobs, const G1* __restrict__ table) {
    const VmsmJob& J = jobs.j[blockIdx.y];
    const uint32_t slot = blockIdx.x * VACC_THREADS + threadIdx.x;
    if (slot >= J.item_start[VNB]) return;
    const uint32_t bucket = J.item_bucket[slot];
    const uint32_t first = J.item_start[bucket], np = J.item_start[bucket + 1] - first, part = slot - first;
    const uint32_t lo = J.starts[bucket], len = J.starts[bucket + 1] - lo;
    const uint32_t b0 = lo + (uint32_t)(((uint64_t)len * part) / np);
    const uint32_t b1 = lo + (uint32_t)(((uint64_t)len * (part + 1)) / np);
    const Fp beta = Fp::from_limbs(FP_BETA_A);
    G1 acc = g1_inf();
#pragma unroll 1
    for (uint32_t k = b0; k < b1; k++) {
        const uint32_t v = J.entries[k];
        G1 t = vload_g1(table + (v & 0x3fffffffu));
        if (v & 0x40000000u) t.x = mul(t.x, beta);
        if (v >> 31) t.y = neg(t.y);
        g1_add_to(acc, t);
    }
    vstore_g1(J.partial + slot, acc);
}

// One CTA per bucket folds the bucket's partial sums pairwise, in place: round after round the upper half
// of the list is added onto the lower half, one cooperative addition (g1_quad.cuh) per quad of lanes.
__global__ void __launch_bounds__(VCOMB_THREADS) vmsm_combine_kernel(const __grid_constant__ VmsmJobs jobs) {
    __shared__ QuadScratch sc[VCOMB_THREADS / 4];
    const VmsmJob& J = jobs.j[blockIdx.y];
    const int t = threadIdx.x, quad = t >> 2, bucket = blockIdx.x;
    const uint32_t it0 = J.item_start[bucket];
    uint32_t m = J.item_start[bucket + 1] - it0;
    G1* base = J.partial + it0;
    while (m > 1) {
        const uint32_t half = (m + 1) >> 1, pairs = m - half;
        for (uint32_t i = quad; i < pairs; i += VCOMB_THREADS / 4) g1_add_quad(base + i, base + i, base + i + half, &sc[quad]);
        __syncthreads();
        m = half;
    }
    if (t < 12) reinterpret_cast<uint4*>(J.combined + bucket)[t] = reinterpret_cast<const uint4*>(base)[t];
}

// sum_b (b+1) B_b = sum_b S_b with S_b = sum_{b' >= b} B_b': suffix scan (ping-pong between `combined`
// and `scan_tmp`), then a tree, one quad per addition.
constexpr int VRED_THREADS = 4 * VNB;
__global__ void __launch_bounds__(VRED_THREADS) vmsm_reduce_kernel(G1* __restrict__ out2, const __grid_constant__ VmsmJobs jobs) {
    extern __shared__ __align__(16) unsigned char vred_smem[];
    QuadScratch* sc = reinterpret_cast<QuadScratch*>(vred_smem);
    const VmsmJob& J = jobs.j[blockIdx.x];
    const int t = threadIdx.x, b = t >> 2, ql = t & 3;
    G1* cur = J.combined;
    G1* nxt = J.scan_tmp;
#pragma unroll 1
    for (int off = 1; off < VNB; off <<= 1) {
        if (b + off < VNB) {
            g1_add_quad(nxt + b, cur + b, cur + b + off, &sc[b]);
        } else {
            // three 16-byte words per lane: a straight copy of the 192-byte point
            for (int w = ql; w < 12; w += 4) reinterpret_cast<uint4*>(nxt + b)[w] = reinterpret_cast<const uint4*>(cur + b)[w];
        }
        __syncthreads();
        G1* tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
#pragma unroll 1
    for (int s = VNB / 2; s > 0; s >>= 1) {
        if (b < s) g1_add_quad(cur + b, cur + b, cur + b + s, &sc[b]);
        __syncthreads();
    }
    if (t < 12) reinterpret_cast<uint4*>(out2 + blockIdx.x)[t] = reinterpret_cast<const uint4*>(cur)[t];
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static size_t val256(size_t x) { return (x + 255) & ~(size_t)255; }
static uint32_t vmsm_max_items(uint64_t nh) { return (uint32_t)(VNB + (nh * 8 + VCAP - 1) / VCAP); }

size_t vmsm_table_points(uint64_t n) { return (size_t)VMSM_LEVELS * (2 * n + 1); }

__global__ void vmsm_neg_generator_levels_kernel(G1* levels) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    G1 p = g1_from_affine(g1a_neg(g1a_generator()));
    G1 q = g1_mul_bls_x_levels(p, levels, 1);
    (void)g1_mul_bls_x_levels(q, levels + 9, 1);
}
// setup: the 18 table levels of -G1 (the base of the sum r^i y_i term)
int launch_vmsm_generator_levels(Launch& L, G1* levels18) {
    vmsm_neg_generator_levels_kernel<<<1, 1, 0, L.stream>>>(levels18);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}
// table levels of n fixed affine bases (optionally negated): one thread per point, setup time only
__global__ void vmsm_point_levels_kernel(G1* __restrict__ levels, const G1Affine* __restrict__ pts, uint32_t n, bool negate) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine a = pts[i];
    if (g1a_is_inf(a)) {
        const G1 inf = g1_inf();
        for (int j = 0; j < G1_LEVELS; j++) g1_store(levels + (size_t)j * n + i, inf);
        return;
    }
    if (negate) a = g1a_neg(a);
    G1 q = g1_mul_bls_x_levels(g1_from_affine(a), levels + i, n);
    (void)g1_mul_bls_x_levels(q, levels + (size_t)9 * n + i, n);
}
int launch_vmsm_point_levels(Launch& L, G1* levels, const G1Affine* pts, uint32_t n, bool negate) {
    vmsm_point_levels_kernel<<<(n + 31) / 32, 32, 0, L.stream>>>(levels, pts, n, negate);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}
// per call: drop those levels into column 2n of the call's table
int vmsm_place_generator(Launch& L, G1* table, uint64_t n) {
    const size_t npts = 2 * n + 1;
    KZG_CUDA_TRY(cudaMemcpy2DAsync(table + 2 * n, npts * sizeof(G1), L.ctx->g_levels, sizeof(G1), sizeof(G1), VMSM_LEVELS, cudaMemcpyDeviceToDevice, L.stream));
    return RET_OK;
}

size_t vmsm_job_bytes(uint64_t nh) {
    const uint32_t mi = vmsm_max_items(nh);
    return val256(nh * VW * sizeof(uint32_t)) + 2 * val256((VNB + 1) * sizeof(uint32_t)) + val256(mi * sizeof(uint32_t)) + val256((size_t)mi * sizeof(G1)) +
           2 * val256(VNB * sizeof(G1)) + val256(VSORT_CTAS * VNB * sizeof(uint32_t));
}
uint8_t* vmsm_job_carve(VmsmJob& J, uint8_t* ws, const uint32_t* halves, uint64_t nh) {
    J.halves = halves;
    J.nh = (uint32_t)nh;
    J.max_items = vmsm_max_items(nh);
    J.entries = (uint32_t*)ws; ws += val256(nh * VW * sizeof(uint32_t));
    J.starts = (uint32_t*)ws; ws += val256((VNB + 1) * sizeof(uint32_t));
    J.item_start = (uint32_t*)ws; ws += val256((VNB + 1) * sizeof(uint32_t));
    J.item_bucket = (uint32_t*)ws; ws += val256(J.max_items * sizeof(uint32_t));
    J.partial = (G1*)ws; ws += val256((size_t)J.max_items * sizeof(G1));
    J.combined = (G1*)ws; ws += val256(VNB * sizeof(G1));
    J.ctahist = (uint32_t*)ws; ws += val256(VSORT_CTAS * VNB * sizeof(uint32_t));
    J.scan_tmp = (G1*)ws; ws += val256(VNB * sizeof(G1));
    return ws;
}

int launch_vmsm_jobs(Launch& L, G1* out2, const VmsmJobs& jobs, const G1* table, uint32_t npts) {
    vmsm_hist_kernel<<<dim3(VSORT_CTAS, 2), VSORT_THREADS, 0, L.stream>>>(jobs);
    KZG_CUDA_TRY(cudaGetLastError());
    vmsm_scatter_kernel<<<dim3(VSORT_CTAS, 2), VSORT_THREADS, 0, L.stream>>>(jobs, npts);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "vmsm_sort");
    dim3 agrid((jobs.j[1].max_items + VACC_THREADS - 1) / VACC_THREADS, 2);
    vmsm_accumulate_kernel<<<agrid, VACC_THREADS, 0, L.stream>>>(jobs, table);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "vmsm_accumulate");
    vmsm_combine_kernel<<<dim3(VNB, 2), VCOMB_THREADS, 0, L.stream>>>(jobs);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_FUNC_ATTR_PER_DEVICE(vmsm_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(VNB * sizeof(QuadScratch)));
    vmsm_reduce_kernel<<<2, VRED_THREADS, VNB * sizeof(QuadScratch), L.stream>>>(out2, jobs);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "vmsm_reduce");
    return RET_OK;
}

size_t rlc_vmsm_scratch_bytes(uint64_t n) {
    const uint64_t nhB = 4 * (2 * n + 1), nhA = 4 * n;
    return val256(nhB * 8) + val256(n * sizeof(Fr)) + vmsm_job_bytes(nhA) + vmsm_job_bytes(nhB);
}

int launch_rlc_vmsm(Launch& L, G1* out2, const G1* table, const Fr* z, const Fr* y, const uint8_t* digest32, uint64_t first, uint64_t n, void* scratch) {
    Digest8 dg;
    for (int i = 0; i < 8; i++)
        dg.h[i] = ((uint32_t)digest32[4 * i] << 24) | ((uint32_t)digest32[4 * i + 1] << 16) | ((uint32_t)digest32[4 * i + 2] << 8) | (uint32_t)digest32[4 * i + 3];
    if (n == 0 || 2 * n + 1 >= (1ull << 30) / VMSM_LEVELS) return RET_ERROR;
    const uint64_t nhB = 4 * (2 * n + 1), nhA = 4 * n;
    uint8_t* ws = (uint8_t*)scratch;
    uint32_t* hB = (uint32_t*)ws; ws += val256(nhB * 8);
    Fr* ty = (Fr*)ws; ws += val256(n * sizeof(Fr));
    VmsmJobs jobs;
    ws = vmsm_job_carve(jobs.j[0], ws, hB + 8 * n, nhA);  // A: the r^i segment (32 bytes per point), points 0..n-1 (proofs)
    ws = vmsm_job_carve(jobs.j[1], ws, hB, nhB);
    const uint32_t npts = (uint32_t)(2 * n + 1);

    rlc_vmsm_scalars_kernel<<<(unsigned)((n + 63) / 64), 64, 0, L.stream>>>(hB, ty, z, y, dg, (uint32_t)n, first);
    KZG_CUDA_TRY(cudaGetLastError());
    rlc_vmsm_ysum_kernel<<<1, 256, 0, L.stream>>>(hB, ty, (uint32_t)n);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "rlc_scalars");
    return launch_vmsm_jobs(L, out2, jobs, table, npts);
}

}  // namespace kzg
