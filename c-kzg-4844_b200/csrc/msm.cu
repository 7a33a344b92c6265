// Fixed-base multi-scalar multiplication over the 4096-point trusted setup, batched over blobs.
//
// Replaces g1_lincomb_fast (src/common/lincomb.c:65-123) -> blst_p1s_mult_pippenger
// (blst/src/multi_scalar.c:370-434) for the calls whose bases are the Lagrange setup:
// blob_to_kzg_commitment (src/eip4844/eip4844.c:253-281) and the quotient commitment in
// compute_kzg_proof_impl (:484).  Same group element, hence the same 48 output bytes.
//
// B200 design (DESIGN.md "MSM"):
//   * the bases are fixed, so load_trusted_setup precomputes T[j][i] = 2^(11 j) * L_i for the 24
//     signed 11-bit windows (9.4 MB, L2 resident).  All windows of a blob then share ONE set of 1024
//     buckets: 98,304 mixed additions per blob and no doublings (the reference does 26 windows x
//     (4096 adds + 1024-bucket reduction) + 250 doublings);
//   * kernel 1 (msm_sort): one CTA per blob turns the blob's big-endian scalars into signed digits and
//     counting-sorts the (window, point) pairs by bucket in shared memory -> per-bucket lists;
//   * kernel 2 (msm_accumulate): ONE THREAD PER WORK ITEM = (bucket, slice of at most `cap` list
//     entries) walks its slice and folds table points into an XYZZ accumulator held in registers
//     (8M+2S each).  No atomics, no shared-memory buckets, no collisions.  Slicing matters: the top
//     window holds only bits 253..254 of the scalar, so its 4096 digits all land in buckets 0..3 --
//     with one thread per bucket those four threads ran 12x longer than the rest (ncu r01b: 20.5 of 32
//     lanes active, 3.3x spread of issued instructions across SMSPs).  `cap` also shrinks for small
//     batches so that a single blob still spreads over the whole GPU;
//   * kernel 3 (msm_reduce): per blob, sum_k k*B_k with 8-bucket running sums per thread, a small
//     scalar multiplication for the chunk offset and a shared-memory tree.
#include <stdlib.h>

#include "engine.h"
#include "g1_hot.cuh"

namespace kzg {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_scalar_be(uint32_t s[8], const uint8_t* p) {
    // 32 bytes big-endian, 16-byte aligned in our staging buffers
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 hi = __ldg(q), lo = __ldg(q + 1);
    s[7] = __byte_perm(hi.x, 0, 0x0123);
    s[6] = __byte_perm(hi.y, 0, 0x0123);
    s[5] = __byte_perm(hi.z, 0, 0x0123);
    s[4] = __byte_perm(hi.w, 0, 0x0123);
    s[3] = __byte_perm(lo.x, 0, 0x0123);
    s[2] = __byte_perm(lo.y, 0, 0x0123);
    s[1] = __byte_perm(lo.z, 0, 0x0123);
    s[0] = __byte_perm(lo.w, 0, 0x0123);
}
__device__ __forceinline__ void load_scalar_le(uint32_t s[8], const uint8_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
    s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
}

// bits [pos, pos+MSM_C) of the 256-bit little-endian value (zero above bit 255)
__device__ __forceinline__ uint32_t window_bits(const uint32_t s[8], int pos) {
    int w = pos >> 5, sh = pos & 31;
    uint32_t lo = (w < 8) ? s[w] : 0u;
    uint32_t hi = (w + 1 < 8) ? s[w + 1] : 0u;
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> sh) & ((1u << MSM_C) - 1u);
}

// Calls f(window j, bucket b, negative) for every non-zero signed digit of s.
template <class Fn>
__device__ __forceinline__ void for_each_digit(const uint32_t s[8], Fn f) {
    uint32_t carry = 0;
#pragma unroll 1
    for (int j = 0; j < MSM_W; j++) {
        uint32_t d = window_bits(s, j * MSM_C) + carry;
        bool negd = d > (uint32_t)MSM_NB;
        carry = negd ? 1u : 0u;
        uint32_t mag = negd ? ((1u << MSM_C) - d) : d;  // 1..1024 (or 0)
        if (mag != 0) f(j, mag - 1u, negd);
    }
}

__device__ __forceinline__ G1Affine load_affine(const G1Affine* p) {
    G1Affine a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = __ldg(q + i);
    return a;
}
__device__ __forceinline__ G1 load_g1(const G1* p) {
    G1 a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) d[i] = q[i];
    return a;
}
__device__ __forceinline__ void store_g1(G1* p, const G1& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* d = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) q[i] = d[i];
}

// ------------------------------------------------------------------------------------------------
// kernel 1: digits + counting sort by bucket, one CTA (256 threads) per blob
// ------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;

__global__ void __launch_bounds__(SORT_THREADS) msm_sort_kernel(
    uint32_t* __restrict__ entries, uint32_t* __restrict__ starts, uint32_t* __restrict__ item_start, uint16_t* __restrict__ item_bucket,
    uint16_t* __restrict__ item_order, const uint8_t* __restrict__ scalars, bool big_endian, int* __restrict__ bad, uint32_t cap, uint32_t max_items
) {
    __shared__ uint32_t lhist[132];  // work items by slice length (<= cap <= 128)
    __shared__ uint32_t hist[MSM_NB];
    __shared__ uint32_t warp_sums[SORT_THREADS / 32];
    __shared__ int s_bad;
    const int blob = blockIdx.x, tid = threadIdx.x;
    const uint8_t* src = scalars + (size_t)blob * BLOB_BYTES;
    uint32_t* out = entries + (size_t)blob * MSM_ENTRIES;

    for (int i = tid; i < MSM_NB; i += SORT_THREADS) hist[i] = 0;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    // pass 1: histogram (+ canonical check of the wire scalars, src/common/bytes.c:64-70)
    for (int i = tid; i < N_BLOB; i += SORT_THREADS) {
        uint32_t s[8];
        if (big_endian) {
            load_scalar_be(s, src + 32 * i);
            if (limbs_geq<8>(s, FR_MOD)) s_bad = 1;
        } else {
            load_scalar_le(s, src + 32 * i);
        }
        for_each_digit(s, [&](int, uint32_t b, bool) { atomicAdd(&hist[b], 1u); });
    }
    __syncthreads();

    // exclusive scan of the 1024 counters: 4 per thread, warp shuffle scan, then across warps
    uint32_t c0 = hist[4 * tid], c1 = hist[4 * tid + 1], c2 = hist[4 * tid + 2], c3 = hist[4 * tid + 3];
    uint32_t tsum = c0 + c1 + c2 + c3, incl = tsum;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; w++) base += warp_sums[w];
    uint32_t excl = base + incl - tsum;
    __syncthreads();
    hist[4 * tid] = excl;
    hist[4 * tid + 1] = excl + c0;
    hist[4 * tid + 2] = excl + c0 + c1;
    hist[4 * tid + 3] = excl + c0 + c1 + c2;
    uint32_t* st = starts + (size_t)blob * (MSM_NB + 1);
    st[4 * tid] = excl;
    st[4 * tid + 1] = excl + c0;
    st[4 * tid + 2] = excl + c0 + c1;
    st[4 * tid + 3] = excl + c0 + c1 + c2;
    if (tid == SORT_THREADS - 1) st[MSM_NB] = excl + tsum;
    // work items: bucket b is cut into ceil(count_b / cap) slices (at least one); second scan
    {
        uint32_t p0 = c0 ? (c0 + cap - 1) / cap : 1, p1 = c1 ? (c1 + cap - 1) / cap : 1;
        uint32_t p2 = c2 ? (c2 + cap - 1) / cap : 1, p3 = c3 ? (c3 + cap - 1) / cap : 1;
        uint32_t ps = p0 + p1 + p2 + p3, pin = ps;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, pin, o);
            if (lane >= o) pin += v;
        }
        __syncthreads();  // warp_sums reuse
        if (lane == 31) warp_sums[warp] = pin;
        __syncthreads();
        uint32_t pb = 0;
        for (int w = 0; w < warp; w++) pb += warp_sums[w];
        uint32_t pe = pb + pin - ps;
        uint32_t* is = item_start + (size_t)blob * (MSM_NB + 1);
        uint16_t* ib = item_bucket + (size_t)blob * max_items;
        uint32_t q = pe;
        is[4 * tid] = q;
        for (uint32_t k = 0; k < p0; k++) ib[q + k] = (uint16_t)(4 * tid);
        q += p0;
        is[4 * tid + 1] = q;
        for (uint32_t k = 0; k < p1; k++) ib[q + k] = (uint16_t)(4 * tid + 1);
        q += p1;
        is[4 * tid + 2] = q;
        for (uint32_t k = 0; k < p2; k++) ib[q + k] = (uint16_t)(4 * tid + 2);
        q += p2;
        is[4 * tid + 3] = q;
        for (uint32_t k = 0; k < p3; k++) ib[q + k] = (uint16_t)(4 * tid + 3);
        q += p3;
        if (tid == SORT_THREADS - 1) is[MSM_NB] = q;
    }
    for (int i = tid; i < 132; i += SORT_THREADS) lhist[i] = 0;
    __syncthreads();
    // Order the work items by slice length (longest first) so that the 32 lanes of a warp of
    // msm_accumulate walk lists of (nearly) equal length: ncu r01k showed 26.2 of 32 lanes active with
    // items in bucket order (Poisson spread of the bucket sizes).
    {
        const uint32_t* is = item_start + (size_t)blob * (MSM_NB + 1);
        const uint16_t* ib = item_bucket + (size_t)blob * max_items;
        uint16_t* ord = item_order + (size_t)blob * max_items;
        const uint32_t total = is[MSM_NB];
        auto slice_len = [&](uint32_t it) {
            const uint32_t b = ib[it], first = is[b], np = is[b + 1] - first, part = it - first;
            const uint32_t len = st[b + 1] - st[b];
            return (uint32_t)(((uint64_t)len * (part + 1)) / np) - (uint32_t)(((uint64_t)len * part) / np);
        };
        for (uint32_t it = tid; it < total; it += SORT_THREADS) atomicAdd(&lhist[slice_len(it)], 1u);
        __syncthreads();
        if (tid == 0) {  // descending exclusive scan over <= 129 bins
            uint32_t run = 0;
            for (int l = 131; l >= 0; l--) {
                uint32_t c = lhist[l];
                lhist[l] = run;
                run += c;
            }
        }
        __syncthreads();
        for (uint32_t it = tid; it < total; it += SORT_THREADS) ord[atomicAdd(&lhist[slice_len(it)], 1u)] = (uint16_t)it;
    }
    __syncthreads();

    // pass 2: scatter (order inside a bucket is irrelevant: group addition commutes)
    for (int i = tid; i < N_BLOB; i += SORT_THREADS) {
        uint32_t s[8];
        if (big_endian)
            load_scalar_be(s, src + 32 * i);
        else
            load_scalar_le(s, src + 32 * i);
        for_each_digit(s, [&](int j, uint32_t b, bool negd) {
            uint32_t pos = atomicAdd(&hist[b], 1u);
            out[pos] = (uint32_t)(j * N_BLOB + i) | (negd ? 0x80000000u : 0u);
        });
    }
    if (tid == 0 && s_bad && bad) bad[blob] = 1;
}

// ------------------------------------------------------------------------------------------------
// kernel 2: bucket accumulation, one thread per (blob, part, bucket)
// ------------------------------------------------------------------------------------------------
constexpr int ACC_THREADS = 128;

template <int MIN_BLOCKS, bool NL>
__global__ void __launch_bounds__(ACC_THREADS, MIN_BLOCKS) msm_accumulate_kernel(
    G1* __restrict__ partial, const uint32_t* __restrict__ entries, const uint32_t* __restrict__ starts,
    const uint32_t* __restrict__ item_start, const uint16_t* __restrict__ item_bucket, const uint16_t* __restrict__ item_order, const G1Affine* __restrict__ table,
    uint32_t max_items
) {
    const uint32_t slot = blockIdx.x * ACC_THREADS + threadIdx.x;
    const int blob = blockIdx.y;
    const uint32_t* is = item_start + (size_t)blob * (MSM_NB + 1);
    if (slot >= is[MSM_NB]) return;
    const uint32_t item = item_order[(size_t)blob * max_items + slot];  // length-sorted schedule
    const uint32_t bucket = item_bucket[(size_t)blob * max_items + item];
    const uint32_t first = is[bucket], np = is[bucket + 1] - first, part = item - first;
    const uint32_t* st = starts + (size_t)blob * (MSM_NB + 1);
    const uint32_t lo = st[bucket], len = st[bucket + 1] - lo;
    const uint32_t b0 = lo + (uint32_t)(((uint64_t)len * part) / np);
    const uint32_t b1 = lo + (uint32_t)(((uint64_t)len * (part + 1)) / np);
    const uint32_t* e = entries + (size_t)blob * MSM_ENTRIES;

    G1 acc = g1_inf();
#pragma unroll 1
    for (uint32_t k = b0; k < b1; k++) {
        uint32_t v = __ldg(e + k);
        G1Affine a = load_affine(table + (v & 0x7fffffffu));
        if (NL)
            g1_madd_nl(acc, a, (v >> 31) != 0);
        else
            g1_madd(acc, a, (v >> 31) != 0);
    }
    store_g1(partial + (size_t)blob * max_items + item, acc);
}

// ------------------------------------------------------------------------------------------------
// kernel 3: sum_b (b+1) * B_b per blob
// ------------------------------------------------------------------------------------------------
// Small batches use a small `cap`, so a bucket can own many partial sums (the four top-window buckets:
// >100 each at cap 8).  One warp per bucket folds them: lane-strided serial sums + a 5-level tree.
__global__ void __launch_bounds__(128) msm_combine_kernel(G1* __restrict__ combined, const G1* __restrict__ partial, const uint32_t* __restrict__ item_start, uint32_t max_items) {
    __shared__ G1 sh[4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bucket = blockIdx.x * 4 + warp, blob = blockIdx.y;
    const uint32_t* is = item_start + (size_t)blob * (MSM_NB + 1);
    const G1* B = partial + (size_t)blob * max_items;
    G1 acc = g1_inf();
#pragma unroll 1
    for (uint32_t it = is[bucket] + lane; it < is[bucket + 1]; it += 32) {
        G1 b = load_g1(B + it);
        g1_add_to(acc, b);
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
    if (lane == 0) store_g1(combined + (size_t)blob * MSM_NB + bucket, sh[warp][0]);
}

constexpr int RED_THREADS = 128;
constexpr int RED_CHUNK = MSM_NB / RED_THREADS;  // 8 buckets per thread

// item_start == nullptr: `partial` already holds one sum per bucket (msm_combine_kernel ran)
__global__ void __launch_bounds__(RED_THREADS) msm_reduce_kernel(G1* __restrict__ result, const G1* __restrict__ partial, const uint32_t* __restrict__ item_start,
                                                                 uint32_t max_items) {
    __shared__ G1 sh[RED_THREADS];
    const int blob = blockIdx.x, t = threadIdx.x;
    const G1* B = partial + (size_t)blob * max_items;
    const uint32_t* is = item_start ? item_start + (size_t)blob * (MSM_NB + 1) : nullptr;

    // running sums over this thread's 8 buckets, top down:
    //   acc = sum_k (k+1) * B[8t+k],  run = sum_k B[8t+k]
    G1 run = g1_inf(), acc = g1_inf();
#pragma unroll 1
    for (int k = RED_CHUNK - 1; k >= 0; k--) {
        const uint32_t bk = (uint32_t)(t * RED_CHUNK + k);
        const uint32_t it0 = is ? is[bk] : bk, it1 = is ? is[bk + 1] : bk + 1;
#pragma unroll 1
        for (uint32_t it = it0; it < it1; it++) {
            G1 b = load_g1(B + it);
            g1_add_to(run, b);
        }
        g1_add_to(acc, run);
    }
    // bucket index b = 8t + k carries weight b + 1 = 8t + (k + 1): add [8t] * run
    if (t != 0) {
        G1 m = g1_inf();
        const uint32_t w = (uint32_t)(t * RED_CHUNK);
#pragma unroll 1
        for (int bit = 31 - __clz(w); bit >= 0; bit--) {
            g1_dbl_to(m);
            if ((w >> bit) & 1u) g1_add_to(m, run);
        }
        g1_add_to(acc, m);
    }
    sh[t] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) {
            G1 x = sh[t], y = sh[t + s];
            g1_add_to(x, y);
            sh[t] = x;
        }
        __syncthreads();
    }
    if (t == 0) store_g1(result + blob, sh[0]);
}

// ------------------------------------------------------------------------------------------------
// point -> 48 bytes
// ------------------------------------------------------------------------------------------------
__global__ void g1_compress_kernel(uint8_t* __restrict__ out48, const G1* __restrict__ pts, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1 p = load_g1(pts + i);
    G1Affine a = g1_to_affine(p);
    uint8_t buf[48];
    g1a_compress(buf, a);
    for (int k = 0; k < 48; k++) out48[i * 48 + k] = buf[k];
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static uint32_t msm_max_items(int cap) {
    uint32_t m = MSM_NB + (MSM_ENTRIES + cap - 1) / cap;  // every bucket one slice + at most one extra per `cap` entries
    return (m + ACC_THREADS - 1) / ACC_THREADS * ACC_THREADS;
}

// `cap` = most list entries one thread folds
size_t msm_workspace_bytes(uint64_t n, int cap) {
    uint32_t mi = msm_max_items(cap);
    return align256(n * MSM_ENTRIES * sizeof(uint32_t)) + 2 * align256(n * (MSM_NB + 1) * sizeof(uint32_t)) + 2 * align256(n * mi * sizeof(uint16_t)) +
           align256(n * mi * sizeof(G1)) + (cap < 128 ? align256(n * MSM_NB * sizeof(G1)) : 0);
}

int msm_pick_parts(uint64_t n) {
    // large batches: 128 entries per thread (mean bucket ~96); small ones: slice finer so that the
    // blob still covers 148 SMs x 4 CTAs of 128 threads
    int cap = 128;
    while (cap > 8 && n * (MSM_NB + MSM_ENTRIES / cap) < 148ull * 512ull) cap /= 2;
    return cap;
}

int launch_msm(Launch& L, G1* result, const uint8_t* scalars, bool big_endian_bytes, uint64_t n, const G1Affine* table, int* d_bad, void* workspace, int cap) {
    if (n == 0) return RET_OK;
    const uint32_t mi = msm_max_items(cap);
    uint8_t* ws = (uint8_t*)workspace;
    uint32_t* entries = (uint32_t*)ws;
    ws += align256(n * MSM_ENTRIES * sizeof(uint32_t));
    uint32_t* starts = (uint32_t*)ws;
    ws += align256(n * (MSM_NB + 1) * sizeof(uint32_t));
    uint32_t* item_start = (uint32_t*)ws;
    ws += align256(n * (MSM_NB + 1) * sizeof(uint32_t));
    uint16_t* item_bucket = (uint16_t*)ws;
    ws += align256(n * mi * sizeof(uint16_t));
    uint16_t* item_order = (uint16_t*)ws;
    ws += align256(n * mi * sizeof(uint16_t));
    G1* partial = (G1*)ws;

    msm_sort_kernel<<<(unsigned)n, SORT_THREADS, 0, L.stream>>>(entries, starts, item_start, item_bucket, item_order, scalars, big_endian_bytes, d_bad, (uint32_t)cap, mi);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "msm_sort");
    dim3 grid(mi / ACC_THREADS, (unsigned)n);
    static const int variant = getenv("CKZG_B200_ACC_VARIANT") ? atoi(getenv("CKZG_B200_ACC_VARIANT")) : 14;  // r01c probe: 14 (noinline multiplier, 128 regs) fastest
    if (variant == 13)
        msm_accumulate_kernel<3, true><<<grid, ACC_THREADS, 0, L.stream>>>(partial, entries, starts, item_start, item_bucket, item_order, table, mi);
    else if (variant == 14)
        msm_accumulate_kernel<4, true><<<grid, ACC_THREADS, 0, L.stream>>>(partial, entries, starts, item_start, item_bucket, item_order, table, mi);
    else if (variant == 4)
        msm_accumulate_kernel<4, false><<<grid, ACC_THREADS, 0, L.stream>>>(partial, entries, starts, item_start, item_bucket, item_order, table, mi);
    else
        msm_accumulate_kernel<3, false><<<grid, ACC_THREADS, 0, L.stream>>>(partial, entries, starts, item_start, item_bucket, item_order, table, mi);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "msm_accumulate");
    if (cap < 128) {
        G1* combined = (G1*)((uint8_t*)partial + align256(n * mi * sizeof(G1)));
        dim3 cgrid(MSM_NB / 4, (unsigned)n);
        msm_combine_kernel<<<cgrid, 128, 0, L.stream>>>(combined, partial, item_start, mi);
        KZG_CUDA_TRY(cudaGetLastError());
        msm_reduce_kernel<<<(unsigned)n, RED_THREADS, 0, L.stream>>>(result, combined, nullptr, MSM_NB);
        KZG_CUDA_TRY(cudaGetLastError());
        L.count(2, "msm_reduce");
    } else {
        msm_reduce_kernel<<<(unsigned)n, RED_THREADS, 0, L.stream>>>(result, partial, item_start, mi);
        KZG_CUDA_TRY(cudaGetLastError());
        L.count(1, "msm_reduce");
    }
    return RET_OK;
}

int launch_g1_compress(Launch& L, uint8_t* out48, const G1* pts, uint64_t n) {
    if (n == 0) return RET_OK;
    g1_compress_kernel<<<(unsigned)((n + 63) / 64), 64, 0, L.stream>>>(out48, pts, n);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "g1_compress");
    return RET_OK;
}

}  // namespace kzg
