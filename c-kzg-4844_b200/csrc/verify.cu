// Per-blob stage and random-linear-combination stage of the EIP-4844 provers / verifiers.
//
// Replaces, for the GPU engine (all paths relative to the reference tree):
//   compute_challenge ......................... src/eip4844/eip4844.c:147-178
//   blob_to_polynomial / bytes_to_bls_field ... src/eip4844/blob.c:31, src/common/bytes.c:64
//   evaluate_polynomial_in_evaluation_form .... src/eip4844/eip4844.c:192-240 (+ fr_batch_inv :80)
//   compute_kzg_proof_impl (quotient) ......... src/eip4844/eip4844.c:417-494
//   compute_r_powers_for_verify_kzg_proof_batch src/eip4844/eip4844.c:597-680
//   verify_kzg_proof_batch (lincombs) ......... src/eip4844/eip4844.c:697-765
//
// Layout: one CTA per blob for the polynomial work (coalesced 32-byte element loads, the 4096
// denominators inverted with ONE field inversion per blob through a block-wide product scan), one
// thread per blob for the inherently serial 131 KB SHA-256, one thread per scalar multiplication in
// the linear combinations.
#include <stdlib.h>
#include <string.h>

#define KZG_FP_MUL_OUTLINE 1
#include "g1_glv.cuh"
#include "g1_quad.cuh"
#include "sha256.cuh"
#include "verify.h"

namespace kzg {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ void load_fr_be(uint32_t s[8], const uint8_t* p) {  // 16-byte aligned
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 hi = __ldg(q), lo = __ldg(q + 1);
    s[7] = bswap32(hi.x); s[6] = bswap32(hi.y); s[5] = bswap32(hi.z); s[4] = bswap32(hi.w);
    s[3] = bswap32(lo.x); s[2] = bswap32(lo.y); s[1] = bswap32(lo.z); s[0] = bswap32(lo.w);
}
__device__ __forceinline__ void store_fr_be(uint8_t* p, const Fr& a) {  // 16-byte aligned
    uint32_t t[8];
    from_mont<FrTag>(t, a);
    uint4 hi = make_uint4(bswap32(t[7]), bswap32(t[6]), bswap32(t[5]), bswap32(t[4]));
    uint4 lo = make_uint4(bswap32(t[3]), bswap32(t[2]), bswap32(t[1]), bswap32(t[0]));
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = hi;
    q[1] = lo;
}
__device__ __forceinline__ Fr load_fr(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// digest (8 words) -> field element: hash_to_bls_field (src/common/bytes.c:123)
__device__ __forceinline__ Fr fr_from_digest(const uint32_t h[8]) {
    uint32_t t[8], s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = h[7 - i];
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        uint32_t bw = limbs_sub<8>(s, t, FR_MOD);
        if (!bw) {
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = s[i];
        }
    }
    return to_mont<FrTag>(t);
}

// ------------------------------------------------------------------------------------------------
// compute_challenge: 2050 dependent SHA-256 compressions per blob -- a pure latency chain.  Two warps
// split it: warp 1 ("schedule") loads the next 64 message bytes of its lane's blob, expands the 64
// schedule words and pre-adds the round constants into a shared-memory buffer; warp 0 ("rounds") runs
// only the 64-round dependency chain of the previous block out of the other buffer.  The two warps sit
// on different SM sub-partitions and meet at one barrier per block.  (One thread doing both issued
// ~1000 instructions per block at IPC 0.26: 4.0 ms per batch whatever its size.)
// ------------------------------------------------------------------------------------------------
constexpr int CH_BLOCKS = 2050;  // (16 + 16 + 131072 + 48 + 1 + 8 bytes) padded to 64-byte blocks

// Placement probe (tools/gpu_probe.py "placement"): when armed, lane 0 of every warp of the stage-1
// kernels records which SM and which hardware warp slot it runs on -- the measurement behind the
// arrangement of those kernels (latency-bound warps that share a sub-partition slow each other down).
__device__ uint32_t* g_place_buf = nullptr;  // [0] = count, [1] = capacity, then records
__device__ __forceinline__ void place_record(uint32_t kernel_id) {
    uint32_t* buf = g_place_buf;
    if (buf == nullptr || (threadIdx.x & 31) != 0) return;
    uint32_t smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    const uint32_t k = atomicAdd(buf, 1u);
    if (k < buf[1]) buf[2 + k] = (kernel_id << 28) | ((threadIdx.x >> 5) << 24) | (smid << 8) | (warpid & 0xffu);
}
// Device-side timeline probe (tools/e2e_probe.py --timers): when armed, thread 0 of CTA 0 of the stage-1 kernels
// stores (id, low 32 bits of %globaltimer) at kernel start (id) and end (id | 0x80) -- when did a kernel really run,
// as opposed to when the events around it fired.
__device__ uint32_t* g_time_buf = nullptr;  // [0] = count, [1] = capacity, then (id, ns) pairs
__device__ __forceinline__ void time_record(uint32_t id) {
    uint32_t* buf = g_time_buf;
    if (buf == nullptr || threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const uint32_t k = atomicAdd(buf, 1u);
    if (k < buf[1]) {
        buf[2 + 2 * k] = id;
        buf[3 + 2 * k] = (uint32_t)t;
    }
}
// CKZG_B200_SHA_ADDS=fma selects the variant with every addition of a round on the FMA pipe (A/B measurements).
// Measured on B200 (profiles/R2_summary.md, R2e): hash+validate 2.28 ms with ptxas' own split (13 ALU + 2 IMAD.IADD
// per round) against 2.50 ms with 10 ALU + 8 IMAD -- integer multiply-adds issue at a quarter warp per clock here
// (the same 32 lanes/clk/SM the Montgomery multiplier measures), so moving a two-cycle ALU addition to the FMA pipe
// costs four cycles there.  The default stays ptxas' split.
static bool sha_fma_adds() {
    static const bool on = getenv("CKZG_B200_SHA_ADDS") && strcmp(getenv("CKZG_B200_SHA_ADDS"), "fma") == 0;
    return on;
}
__device__ __forceinline__ void pair_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

// message block `k` of the transcript "FSBLOBVERIFY_V1_" || u64be(0) || u64be(4096) || blob || commitment
__device__ __forceinline__ void challenge_message_block(uint32_t w[16], int k, const uint4* __restrict__ blob, const uint8_t* __restrict__ cm) {
    if (k == 0) {
        w[0] = 0x4653424cu; w[1] = 0x4f425645u; w[2] = 0x52494659u; w[3] = 0x5f56315fu;  // "FSBL" "OBVE" "RIFY" "_V1_"
        w[4] = 0; w[5] = 0; w[6] = 0; w[7] = 4096;
        uint4 a = __ldg(blob), b = __ldg(blob + 1);
        w[8] = bswap32(a.x); w[9] = bswap32(a.y); w[10] = bswap32(a.z); w[11] = bswap32(a.w);
        w[12] = bswap32(b.x); w[13] = bswap32(b.y); w[14] = bswap32(b.z); w[15] = bswap32(b.w);
    } else if (k < 2048) {
        const uint4* p = blob + (4 * k - 2);
        uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
        w[0] = bswap32(v0.x); w[1] = bswap32(v0.y); w[2] = bswap32(v0.z); w[3] = bswap32(v0.w);
        w[4] = bswap32(v1.x); w[5] = bswap32(v1.y); w[6] = bswap32(v1.z); w[7] = bswap32(v1.w);
        w[8] = bswap32(v2.x); w[9] = bswap32(v2.y); w[10] = bswap32(v2.z); w[11] = bswap32(v2.w);
        w[12] = bswap32(v3.x); w[13] = bswap32(v3.y); w[14] = bswap32(v3.z); w[15] = bswap32(v3.w);
    } else if (k == 2048) {
        uint4 v0 = __ldg(blob + 8190), v1 = __ldg(blob + 8191);
        w[0] = bswap32(v0.x); w[1] = bswap32(v0.y); w[2] = bswap32(v0.z); w[3] = bswap32(v0.w);
        w[4] = bswap32(v1.x); w[5] = bswap32(v1.y); w[6] = bswap32(v1.z); w[7] = bswap32(v1.w);
#pragma unroll
        for (int j = 0; j < 8; j++) w[8 + j] = be32(cm + 4 * j);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = be32(cm + 32 + 4 * j);
        w[4] = 0x80000000u;
#pragma unroll
        for (int j = 5; j < 15; j++) w[j] = 0;
        w[15] = 131152u * 8u;
    }
}

// Runs on warps 0 and 1 of the CTA (they meet at named barrier 1, so other warps of the CTA are free
// to do something else).
// `lanes` (<= 32) blobs per CTA: the lanes above it leave at once (a partially filled warp still costs
// the full issue slots, so this only helps to spread a small batch over more SMs).
//
// FMA_ADDS (experiment, off by default: see sha_fma_adds): the rounds warp is bound by the 16-lane ALU pipe, not by
// latency -- every warp instruction holds its pipe for two cycles and a round is 6 SHF + 4 LOP3 + 3 IADD3 on the ALU
// pipe (13 x 2 = 26 of the 31 cycles measured per round) against 2 IMAD.IADD on the FMA pipe.  With FMA_ADDS every
// addition of the round is written as x * one + y with
// a run-time `one` (%nsmid clamped to 1), which ptxas cannot turn back into IADD3: 10 ALU + 8 FMA instructions per
// round, and the e-recurrence is still three dependent instructions (SHF -> LOP3 -> IMAD).
__device__ __forceinline__ uint32_t fma_add(uint32_t x, uint32_t one, uint32_t y) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(y));
    return r;
}
//
// Block range [k0, k1) with the SHA state carried in `states` (8 words per blob): a blob that arrives over PCIe in
// column pieces is hashed piece by piece, each launch continuing where the last one stopped (api_verify.cu: the last
// chunk of a host-pointer batch), so that only the final piece's 256 blocks remain when its last byte lands.
template <bool FMA_ADDS>
__device__ __forceinline__ void challenge_warp_pair(uint32_t (*kw)[64][32], Fr* __restrict__ z_out, uint8_t* __restrict__ zy, const uint8_t* __restrict__ blobs,
                                                    const uint8_t* __restrict__ commitments, uint64_t n, int lanes, int k0 = 0, int k1 = CH_BLOCKS,
                                                    uint32_t* __restrict__ states = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= lanes) return;
    const uint64_t i = (uint64_t)blockIdx.x * lanes + lane;
    const bool live = i < n;
    const uint64_t ii = live ? i : 0;  // dead lanes shadow blob 0 (reads only) so every thread reaches every barrier
    const uint4* blob = reinterpret_cast<const uint4*>(blobs + ii * BLOB_BYTES);
    const uint8_t* cm = commitments + ii * 48;
    Sha256 st;
    sha256_init(st);
    if (k0 > 0 && warp == 0) {
#pragma unroll
        for (int j = 0; j < 8; j++) st.h[j] = states[ii * 8 + j];
    }
    uint4 nx0, nx1, nx2, nx3;  // schedule warp: the NEXT block's 64 message bytes, in flight during this block's expansion
    nx0 = nx1 = nx2 = nx3 = make_uint4(0, 0, 0, 0);
#pragma unroll 1
    for (int k = k0; k <= k1; k++) {
        if (warp == 1) {
            if (k < k1) {
                uint32_t w[16];
                if (k > k0 && k >= 1 && k < 2048) {
                    w[0] = bswap32(nx0.x); w[1] = bswap32(nx0.y); w[2] = bswap32(nx0.z); w[3] = bswap32(nx0.w);
                    w[4] = bswap32(nx1.x); w[5] = bswap32(nx1.y); w[6] = bswap32(nx1.z); w[7] = bswap32(nx1.w);
                    w[8] = bswap32(nx2.x); w[9] = bswap32(nx2.y); w[10] = bswap32(nx2.z); w[11] = bswap32(nx2.w);
                    w[12] = bswap32(nx3.x); w[13] = bswap32(nx3.y); w[14] = bswap32(nx3.z); w[15] = bswap32(nx3.w);
                } else {
                    challenge_message_block(w, k, blob, cm);
                }
                if (k + 1 < 2048 && k + 1 < k1) {  // a global load costs about as much as half an expansion: issue it a block ahead
                    const uint4* p = blob + (4 * (k + 1) - 2);
                    nx0 = __ldg(p); nx1 = __ldg(p + 1); nx2 = __ldg(p + 2); nx3 = __ldg(p + 3);
                }
                // 16 rounds of code, run four times: the loop stays a few KB so that the four warps of
                // an SM (hash pair + two chain warps) fit the instruction cache together (ncu r01t: the
                // fully unrolled pair was 30 KB, icc hit rate 79 %, no_instruction 0.7 stalls per issue)
                uint32_t(*dst)[32] = kw[k & 1];
#pragma unroll
                for (int t = 0; t < 16; t++) dst[t][lane] = w[t] + SHA256_K[t];
#pragma unroll 1
                for (int tt = 16; tt < 64; tt += 16) {
#pragma unroll
                    for (int t = 0; t < 16; t++) {
                        uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                        uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                        uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                        w[t] = w[t] + s0 + w[(t + 9) & 15] + s1;
                        dst[tt + t][lane] = w[t] + SHA256_K[tt + t];
                    }
                }
            }
        } else if (k > k0) {
            const uint32_t(*src)[32] = kw[(k - 1) & 1];
            uint32_t a = st.h[0], b = st.h[1], c = st.h[2], d = st.h[3], e = st.h[4], f = st.h[5], g = st.h[6], h = st.h[7];
            uint32_t one = 1u;
            if (FMA_ADDS) {  // a value neither the front end nor ptxas can fold: min(number of SMs, 1)
                asm volatile("mov.u32 %0, %%nsmid;" : "=r"(one));
                one = one < 1u ? one : 1u;
            }
#pragma unroll 1
            for (int tt = 0; tt < 64; tt += 16) {
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    if (FMA_ADDS) {
                        const uint32_t kwv = src[tt + t][lane];
                        const uint32_t ch = (e & f) ^ (~e & g);
                        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
                        const uint32_t pa = fma_add(h, one, kwv);
                        const uint32_t pe = fma_add(pa, one, d);
                        const uint32_t u = fma_add(ch, one, pe);       // everything of e' except Sigma1
                        const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
                        const uint32_t q1 = fma_add(S0, one, pa);
                        const uint32_t q2 = fma_add(mj, one, q1);
                        const uint32_t w2 = fma_add(ch, one, q2);      // everything of a' except Sigma1
                        const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
                        const uint32_t en = fma_add(S1, one, u);
                        const uint32_t an = fma_add(S1, one, w2);
                        h = g; g = f; f = e; e = en;
                        d = c; c = b; b = a; a = an;
                        continue;
                    }
                    // the recurrence through e is the critical path: everything that does not depend on the
                    // current e or a (h, d, K+W) is summed first, so e' is ONE three-input add behind
                    // Sigma1/Ch -- rotate, xor, add: three dependent instructions per round instead of five
                    uint32_t pa = h + src[tt + t][lane];
                    uint32_t pe = pa + d;
                    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
                    uint32_t ch = (e & f) ^ (~e & g);
                    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
                    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
                    uint32_t qa = pa + S0 + mj;
                    uint32_t en = pe + S1 + ch;
                    uint32_t an = qa + S1 + ch;
                    h = g; g = f; f = e; e = en;
                    d = c; c = b; b = a; a = an;
                }
            }
            st.h[0] += a; st.h[1] += b; st.h[2] += c; st.h[3] += d;
            st.h[4] += e; st.h[5] += f; st.h[6] += g; st.h[7] += h;
        }
        pair_barrier();
    }
    if (warp == 0 && live) {
        if (k1 < CH_BLOCKS) {
#pragma unroll
            for (int j = 0; j < 8; j++) states[i * 8 + j] = st.h[j];
        } else {
            Fr z = fr_from_digest(st.h);
            z_out[i] = z;
            store_fr_be(zy + i * 64, z);
        }
    }
}

template <bool FMA_ADDS>
__global__ void __launch_bounds__(64) blob_challenge_kernel(Fr* __restrict__ z_out, uint8_t* __restrict__ zy, const uint8_t* __restrict__ blobs,
                                                          const uint8_t* __restrict__ commitments, uint64_t n, int lanes, int k0, int k1, uint32_t* __restrict__ states) {
    __shared__ uint32_t kw[2][64][32];  // [buffer][round][lane] = K[t] + W[t]
    place_record(1);
    time_record(1u | ((uint32_t)k0 << 8));
    challenge_warp_pair<FMA_ADDS>(kw, z_out, zy, blobs, commitments, n, lanes, k0, k1, states);
    time_record(0x81u | ((uint32_t)k0 << 8));
}

__global__ void z_from_bytes_kernel(Fr* z_out, uint8_t* zy, const uint8_t* z_bytes, uint64_t n, int* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[32];
    for (int k = 0; k < 32; k++) buf[k] = z_bytes[32 * i + k];
    Fr z;
    if (!fr_from_be_checked(z, buf)) bad[i] = 1;
    z_out[i] = z;
    if (zy)
        for (int k = 0; k < 32; k++) zy[64 * i + k] = buf[k];
}

__global__ void fr_from_bytes_kernel(Fr* out, const uint8_t* bytes, uint64_t stride, uint64_t n, int* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[32];
    for (int k = 0; k < 32; k++) buf[k] = bytes[stride * i + k];
    Fr v;
    if (!fr_from_be_checked(v, buf) && bad) bad[i] = 1;
    out[i] = v;
}

// ------------------------------------------------------------------------------------------------
// barycentric evaluation: one CTA of 256 threads per blob, 16 elements per thread
// ------------------------------------------------------------------------------------------------
constexpr int EV_THREADS = 256;
constexpr int EV_PER = N_BLOB / EV_THREADS;  // 16

// total[blob] = prod_i (z - w_i) over the 4096 evaluation points (a factor that is zero -- z in the
// domain -- is replaced by 1, exactly as evaluate_kernel does)
__global__ void __launch_bounds__(EV_THREADS) evaluate_products_kernel(Fr* __restrict__ total, const Fr* __restrict__ z_in, const Fr* __restrict__ roots_brp) {
    __shared__ Fr sh[EV_THREADS];
    const int blob = blockIdx.x, t = threadIdx.x;
    const Fr z = z_in[blob];
    Fr prod = Fr::one();
#pragma unroll 1
    for (int k = 0; k < EV_PER; k++) {
        Fr d = sub(z, load_fr(roots_brp + t + EV_THREADS * k));
        if (is_zero(d)) d = Fr::one();
        prod = mul(prod, d);
    }
    sh[t] = prod;
    __syncthreads();
#pragma unroll 1
    for (int s = EV_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) sh[t] = mul(sh[t], sh[t + s]);
        __syncthreads();
    }
    if (t == 0) total[blob] = sh[0];
}
__global__ void evaluate_invert_kernel(Fr* __restrict__ total, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) total[i] = fr_inv(total[i]);
}

__global__ void __launch_bounds__(EV_THREADS) evaluate_kernel(Fr* __restrict__ y_out, uint8_t* __restrict__ zy, Fr* __restrict__ inv_out, int* __restrict__ m_out,
                                                             const uint8_t* __restrict__ blobs, const Fr* __restrict__ z_in, const Fr* __restrict__ roots_brp,
                                                             const Fr* __restrict__ total_inv, int* __restrict__ bad, int bad_stride) {
    __shared__ Fr sh_a[EV_THREADS];  // prefix scan / reduction scratch
    __shared__ Fr sh_b[EV_THREADS];  // suffix scan
    __shared__ Fr sh_bcast;
    __shared__ int s_m, s_bad;
    const int blob = blockIdx.x, t = threadIdx.x;
    const uint8_t* src = blobs + (size_t)blob * BLOB_BYTES;
    const Fr z = z_in[blob];
    if (t == 0) {
        s_m = -1;
        s_bad = 0;
    }
    __syncthreads();

    // pass 1: running product of this thread's denominators z - w_i (the blob is not touched yet)
    Fr pre[EV_PER];
    Fr prod = Fr::one();
#pragma unroll 1
    for (int k = 0; k < EV_PER; k++) {
        int i = t + EV_THREADS * k;
        Fr d = sub(z, load_fr(roots_brp + i));
        if (is_zero(d)) {  // z is the i-th evaluation point (eip4844.c:213)
            s_m = i;
            d = Fr::one();
        }
        pre[k] = prod;
        prod = mul(prod, d);
    }
    // block-wide exclusive prefix and suffix products of the 256 per-thread products
    sh_a[t] = prod;
    sh_b[t] = prod;
    __syncthreads();
#pragma unroll 1
    for (int off = 1; off < EV_THREADS; off <<= 1) {
        Fr pa = sh_a[t], pb = sh_b[t];
        bool ha = t >= off, hb = t + off < EV_THREADS;
        Fr xa = ha ? sh_a[t - off] : pa;
        Fr xb = hb ? sh_b[t + off] : pb;
        __syncthreads();
        if (ha) sh_a[t] = mul(pa, xa);
        if (hb) sh_b[t] = mul(pb, xb);
        __syncthreads();
    }
    // inclusive scans done: sh_a[t] = prod_0..t, sh_b[t] = prod_t..255
    // the one inversion of this blob was done by evaluate_invert_kernel (one thread per blob: a
    // 380-product dependent chain must not hold 255 other threads at a barrier -- ncu r01g)
    if (t == 0) sh_bcast = total_inv[blob];
    __syncthreads();
    Fr acc = sh_bcast;
    if (t > 0) acc = mul(acc, sh_a[t - 1]);
    if (t < EV_THREADS - 1) acc = mul(acc, sh_b[t + 1]);
    // acc = 1 / (this thread's product)
    __syncthreads();

    // pass 2: unwind -> 1/(z - w_i); accumulate p_i * w_i / (z - w_i)
    Fr sum = Fr::zero();
    const int m = s_m;
#pragma unroll 1
    for (int k = EV_PER - 1; k >= 0; k--) {
        int i = t + EV_THREADS * k;
        Fr w = load_fr(roots_brp + i);
        Fr d = sub(z, w);
        if (i == m) d = Fr::one();
        Fr inv = mul(acc, pre[k]);
        acc = mul(acc, d);
        if (inv_out) inv_out[(size_t)blob * N_BLOB + i] = inv;
        // The blob is read exactly once, here.  The big-endian integer p_i is used as it stands, i.e. as
        // the Montgomery form of p_i / R: the sum comes out scaled by 1/R and one product with R^2 at the
        // end puts it right -- 4096 to-Montgomery products per blob saved, and no per-thread array of
        // converted elements to keep in local memory between the passes (ncu r01k: 437 KB/blob of it).
        Fr praw;
        load_fr_be(praw.l, src + 32 * i);
        if (limbs_geq<8>(praw.l, FR_MOD)) s_bad = 1;  // bytes_to_bls_field, bytes.c:67
        sum = add(sum, mul(mul(praw, w), inv));
    }
    sh_a[t] = sum;
    __syncthreads();
#pragma unroll 1
    for (int s = EV_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) sh_a[t] = add(sh_a[t], sh_a[t + s]);
        __syncthreads();
    }
    if (t == 0) {
        Fr y;
        if (m >= 0) {
            uint32_t s[8];
            load_fr_be(s, src + 32 * m);
            y = to_mont<FrTag>(s);
        } else {
            // sum * (z^4096 - 1) / 4096
            Fr zp = z;
#pragma unroll 1
            for (int k = 0; k < 12; k++) zp = sqr(zp);
            y = mul(mul(sh_a[0], Fr::from_limbs(FR_INV_4096)), sub(zp, Fr::one()));
            y = mul(y, Fr::from_limbs(FR_R2));  // undo the 1/R carried by the raw blob elements
        }
        y_out[blob] = y;
        if (zy) store_fr_be(zy + (size_t)blob * 64 + 32, y);
        if (m_out) m_out[blob] = m;
        if (s_bad && bad) bad[(size_t)blob * bad_stride] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Inversion-free evaluation for the verifiers (no per-element inverses are needed there).
//
//   p(z) = (z^N - 1)/N * sum_i f_i w_i / (z - w_i)  =  (1/N) * sum_i f_i w_i * prod_{j != i} (z - w_j)
//
// The blob is in bit-reversed order, so leaves 2j, 2j+1 are the points +-w and every node of the
// binary tree over the blob owns the points of one binomial  z^(2m) - c^2 = (z^m - c)(z^m + c).
// With N_node = sum_{i in node} f_i w_i prod_{j in node, j != i} (z - w_j):
//       N_node = N_left (z^m + c) + N_right (z^m - c) = (N_left + N_right) z^m + (N_left - N_right) c,
// c = roots_brp[2 * node index] at every level, and N_root / N = p(z): two products per node, one per
// leaf, ~12.3k products per blob instead of 28.7k (prefix products, one inversion, back-substitution),
// no inversion kernel in the middle, and no special case for z inside the domain -- the identity is
// a polynomial one (evaluate_polynomial_in_evaluation_form's shortcut, eip4844.c:213, returns the
// same value).  Raw big-endian elements are used as Montgomery forms of f/R; the factor comes back
// with the final constant.
// ------------------------------------------------------------------------------------------------
KZG_CONST uint32_t FR_INV4096_R2[8] = {0x5fdf3f2au, 0xc90c999eu, 0x48481b52u, 0x4997a3e2u, 0x950dfcb5u, 0x8201d923u, 0x8175d40cu, 0x19e61b5eu};  // R^2 / 4096 mod r

__global__ void evaluate_zpow_kernel(Fr* __restrict__ zpow, const Fr* __restrict__ z, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr p = z[i];
    zpow[i * 12] = p;
#pragma unroll 1
    for (int k = 1; k < 12; k++) {
        p = sqr(p);
        zpow[i * 12 + k] = p;
    }
}

// one out-of-line Fr multiplier for the tree: 62 inlined products made the kernel 222 KB of straight-line
// code (1.59 ms per 4096 blobs, r01v)
static __device__ __noinline__ Fr fr_mul_nl(Fr a, Fr b) { return mul(a, b); }

template <int LOG>
__device__ __forceinline__ Fr evt_subtree(const uint8_t* __restrict__ src, int leaf0, const Fr* Z, const Fr* __restrict__ roots_brp, bool& bad) {
    if constexpr (LOG == 0) {
        Fr praw;
        load_fr_be(praw.l, src + 32 * leaf0);
        if (limbs_geq<8>(praw.l, FR_MOD)) bad = true;  // bytes_to_bls_field, bytes.c:67
        return fr_mul_nl(praw, load_fr(roots_brp + leaf0));
    } else {
        Fr l = evt_subtree<LOG - 1>(src, leaf0, Z, roots_brp, bad);
        Fr r = evt_subtree<LOG - 1>(src, leaf0 + (1 << (LOG - 1)), Z, roots_brp, bad);
        Fr c = load_fr(roots_brp + 2 * (leaf0 >> LOG));
        return add(fr_mul_nl(add(l, r), Z[LOG - 1]), fr_mul_nl(sub(l, r), c));
    }
}

// TREE_LOG leaves-per-thread exponent: 2^TREE_LOG consecutive leaves and the TREE_LOG levels above them stay in one
// thread's registers, the remaining 12 - TREE_LOG levels go through shared memory with half the threads dropping out
// per level.  4 (256 threads) and 5 (128 threads) are both built; CKZG_B200_EVAL_LEAVES selects (A/B, profiles/).
template <int TREE_LOG>
__global__ void __launch_bounds__(N_BLOB >> TREE_LOG) evaluate_tree_kernel(Fr* __restrict__ y_out, uint8_t* __restrict__ zy, const uint8_t* __restrict__ blobs,
                                                                          const Fr* __restrict__ zpow, const Fr* __restrict__ roots_brp, int* __restrict__ bad, int bad_stride) {
    constexpr int THREADS = N_BLOB >> TREE_LOG;
    __shared__ Fr sh[THREADS];
    const int blob = blockIdx.x, t = threadIdx.x;
    time_record(3u);
    const uint8_t* src = blobs + (size_t)blob * BLOB_BYTES;
    const Fr* zp = zpow + (size_t)blob * 12;
    Fr Z[TREE_LOG];
#pragma unroll
    for (int k = 0; k < TREE_LOG; k++) Z[k] = load_fr(zp + k);
    bool isbad = false;
    Fr v = evt_subtree<TREE_LOG>(src, (1 << TREE_LOG) * t, Z, roots_brp, isbad);
    if (isbad && bad) bad[(size_t)blob * bad_stride] = 1;
    sh[t] = v;
    __syncthreads();
#pragma unroll 1
    for (int k = TREE_LOG; k < 12; k++) {
        const int active = THREADS >> (k - TREE_LOG + 1);
        Fr nv;
        if (t < active) {
            Fr l = sh[2 * t], r = sh[2 * t + 1];
            nv = add(fr_mul_nl(add(l, r), load_fr(zp + k)), fr_mul_nl(sub(l, r), load_fr(roots_brp + 2 * t)));
        }
        __syncthreads();
        if (t < active) sh[t] = nv;
        __syncthreads();
    }
    if (t == 0) {
        Fr y = mul(sh[0], Fr::from_limbs(FR_INV4096_R2));
        if (y_out) y_out[blob] = y;
        if (zy) store_fr_be(zy + (size_t)blob * 64 + 32, y);
    }
}

// ------------------------------------------------------------------------------------------------
// quotient polynomial (evaluation form) -> plain little-endian scalars for the MSM
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quotient_kernel(uint8_t* __restrict__ q_out, const uint8_t* __restrict__ blobs, const Fr* __restrict__ y_in,
                                                       const Fr* __restrict__ inv, const int* __restrict__ m_in) {
    const int blob = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[8];
    load_fr_be(s, blobs + (size_t)blob * BLOB_BYTES + 32 * i);
    Fr p = to_mont<FrTag>(s);
    // (p_i - y) / (w_i - z) = (y - p_i) * 1/(z - w_i)
    Fr q = mul(sub(y_in[blob], p), load_fr(inv + (size_t)blob * N_BLOB + i));
    if (i == m_in[blob]) q = Fr::zero();  // fixed up by quotient_in_domain_kernel
    uint32_t o[8];
    from_mont<FrTag>(o, q);
    uint4* dst = reinterpret_cast<uint4*>(q_out + (size_t)blob * BLOB_BYTES + 32 * i);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// z == w_m: q_m = (1/z) * sum_{i != m} (p_i - y) * w_i / (z - w_i)      (eip4844.c:460-481)
__global__ void __launch_bounds__(256) quotient_in_domain_kernel(uint8_t* __restrict__ q_out, const uint8_t* __restrict__ blobs, const Fr* __restrict__ z_in,
                                                                 const Fr* __restrict__ y_in, const Fr* __restrict__ inv, const Fr* __restrict__ roots_brp,
                                                                 const int* __restrict__ m_in) {
    __shared__ Fr sh[256];
    const int blob = blockIdx.x, t = threadIdx.x;
    const int m = m_in[blob];
    if (m < 0) return;
    const Fr y = y_in[blob];
    Fr sum = Fr::zero();
#pragma unroll 1
    for (int k = 0; k < N_BLOB / 256; k++) {
        int i = t + 256 * k;
        if (i == m) continue;
        uint32_t s[8];
        load_fr_be(s, blobs + (size_t)blob * BLOB_BYTES + 32 * i);
        Fr p = to_mont<FrTag>(s);
        Fr term = mul(mul(sub(p, y), load_fr(roots_brp + i)), load_fr(inv + (size_t)blob * N_BLOB + i));
        sum = add(sum, term);
    }
    sh[t] = sum;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) sh[t] = add(sh[t], sh[t + s]);
        __syncthreads();
    }
    if (t == 0) {
        Fr q = mul(sh[0], fr_inv(z_in[blob]));
        uint32_t o[8];
        from_mont<FrTag>(o, q);
        uint32_t* dst = reinterpret_cast<uint32_t*>(q_out + (size_t)blob * BLOB_BYTES + 32 * m);
        for (int j = 0; j < 8; j++) dst[j] = o[j];
    }
}

// ------------------------------------------------------------------------------------------------
// point validation
// ------------------------------------------------------------------------------------------------
__global__ void g1_validate_kernel(G1Affine* __restrict__ out, const uint8_t* __restrict__ bytes, uint64_t n, int* __restrict__ bad, int bad_stride) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[48];
    for (int k = 0; k < 48; k++) buf[k] = bytes[i * 48 + k];
    G1Affine a;
    if (!g1a_validate(a, buf)) {
        bad[i * bad_stride] = 1;
        a = g1a_inf();
    }
    out[i] = a;
}

// ------------------------------------------------------------------------------------------------
// batch challenge r: the transcript hash runs on the host (src/host_sha256.c); the device reduces the
// 32-byte digest mod r (hash_to_bls_field, src/common/bytes.c:123)
// ------------------------------------------------------------------------------------------------
__global__ void r_from_digest_kernel(Fr* r_out, const uint8_t* __restrict__ digest) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t h[8];
    for (int i = 0; i < 8; i++) h[i] = be32(digest + 4 * i);
    *r_out = fr_from_digest(h);
}

// ------------------------------------------------------------------------------------------------
// random linear combination
// ------------------------------------------------------------------------------------------------
// s1[i] = r^(first+i), s2[i] = s1[i] * z_i (plain limbs), ty[i] = s1[i] * y_i (Montgomery)
__global__ void rlc_scalars_kernel(uint32_t* __restrict__ s1, uint32_t* __restrict__ s2, Fr* __restrict__ ty, const Fr* __restrict__ z, const Fr* __restrict__ y,
                                   const Fr* __restrict__ r, bool use_r, uint64_t first, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr p = Fr::one();
    if (use_r) {
        Fr base = *r;
        uint64_t e = first + i;
        while (e) {
            if (e & 1) p = mul(p, base);
            base = sqr(base);
            e >>= 1;
        }
    }
    from_mont<FrTag>(s1 + 8 * i, p);
    from_mont<FrTag>(s2 + 8 * i, mul(p, z[i]));
    ty[i] = mul(p, y[i]);
}

// sum of n field elements -> plain limbs of -(sum) ... kept positive here; the caller subtracts the point
__global__ void __launch_bounds__(256) fr_sum_kernel(uint32_t* out_plain, const Fr* __restrict__ in, uint64_t n) {
    __shared__ Fr sh[256];
    const int t = threadIdx.x;
    Fr s = Fr::zero();
    for (uint64_t i = t; i < n; i += 256) s = add(s, in[i]);
    sh[t] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (t < k) sh[t] = add(sh[t], sh[t + k]);
        __syncthreads();
    }
    if (t == 0) from_mont<FrTag>(out_plain, sh[0]);
}

// thread j < n: U[j] = [s1_j] proof_j; n <= j < 2n: T[j-n] = [s2] proof; 2n <= j < 3n: T[j-n] = [s1] C;
// j == 3n: Y = -[ysum] G1
__global__ void __launch_bounds__(64) rlc_points_kernel(G1* __restrict__ U, G1* __restrict__ T, G1* __restrict__ Y, const G1Affine* __restrict__ cm, const G1Affine* __restrict__ pf,
                                                        const uint32_t* __restrict__ s1, const uint32_t* __restrict__ s2, const uint32_t* __restrict__ ysum, uint64_t n) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > 3 * n) return;
    G1Affine base;
    const uint32_t* k;
    G1* dst;
    if (j < n) {
        base = pf[j]; k = s1 + 8 * j; dst = U + j;
    } else if (j < 2 * n) {
        base = pf[j - n]; k = s2 + 8 * (j - n); dst = T + (j - n);
    } else if (j < 3 * n) {
        base = cm[j - 2 * n]; k = s1 + 8 * (j - 2 * n); dst = T + (j - n);
    } else {
        base = g1a_neg(g1a_generator()); k = ysum; dst = Y;
    }
    uint32_t kk[8];
    for (int q = 0; q < 8; q++) kk[q] = k[q];
    *dst = g1_mul_glv_affine(base, kk);  // scalars are field elements (< r)
}

// tree sum: each block folds up to 128 * 8 inputs into one output
constexpr int SUM_THREADS = 128;
constexpr int SUM_PER = 8;
__global__ void __launch_bounds__(SUM_THREADS) g1_sum_kernel(G1* __restrict__ out, const G1* __restrict__ in, uint64_t n) {
    __shared__ G1 sh[SUM_THREADS];
    const int t = threadIdx.x;
    uint64_t base = (uint64_t)blockIdx.x * SUM_THREADS * SUM_PER;
    G1 acc = g1_inf();
#pragma unroll 1
    for (int k = 0; k < SUM_PER; k++) {
        uint64_t i = base + (uint64_t)k * SUM_THREADS + t;
        if (i < n) {
            G1 p = in[i];
            g1_add_to(acc, p);
        }
    }
    sh[t] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = SUM_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) {
            G1 x = sh[t], y = sh[t + s];
            g1_add_to(x, y);
            sh[t] = x;
        }
        __syncthreads();
    }
    if (t == 0) out[blockIdx.x] = sh[0];
}

// The two sums of the linear combination, folded together: level 1 (grid.y = 0: U, 1: T) turns 256
// inputs into one partial per block, level 2 is one block whose warp 0 finishes A = sum U and whose
// warp 1 finishes B = sum T + Y.  Two launches on the critical path of every verify call instead of the
// five launches and five 192-byte copies of three generic tree sums (0.61 -> ~0.25 ms at n = 4096).
constexpr int FOLD_THREADS = 128;
constexpr int FOLD_IN = 2 * FOLD_THREADS;
__global__ void __launch_bounds__(FOLD_THREADS) rlc_fold_kernel(G1* __restrict__ PU, G1* __restrict__ PT, const G1* __restrict__ U, const G1* __restrict__ T, uint64_t n) {
    __shared__ G1 sh[FOLD_THREADS];
    const int t = threadIdx.x;
    const G1* in = blockIdx.y ? T : U;
    const uint64_t m = blockIdx.y ? 2 * n : n;
    const uint64_t base = (uint64_t)blockIdx.x * FOLD_IN;
    if (base >= m) return;  // whole block: the grid is sized for T
    G1 acc = g1_inf();
    if (base + t < m) acc = in[base + t];
    if (base + FOLD_THREADS + t < m) {
        G1 p = in[base + FOLD_THREADS + t];
        g1_add_to(acc, p);
    }
    sh[t] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = FOLD_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) {
            G1 x = sh[t], y = sh[t + s];
            g1_add_to(x, y);
            sh[t] = x;
        }
        __syncthreads();
    }
    if (t == 0) (blockIdx.y ? PT : PU)[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(64) rlc_final_kernel(G1* __restrict__ out2, const G1* __restrict__ PU, uint64_t nu, const G1* __restrict__ PT, uint64_t nt, const G1* __restrict__ Y) {
    __shared__ G1 sh[64];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const G1* in = w ? PT : PU;
    const uint64_t cnt = w ? nt : nu;
    G1 acc = g1_inf();
    if (w == 1 && lane == 31) acc = *Y;
#pragma unroll 1
    for (uint64_t i = lane; i < cnt; i += 32) {
        G1 p = in[i];
        g1_add_to(acc, p);
    }
    sh[t] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = 16; s > 0; s >>= 1) {
        if (lane < s) {
            G1 x = sh[t], y = sh[t + s];
            g1_add_to(x, y);
            sh[t] = x;
        }
        __syncthreads();
    }
    if (lane == 0) out2[w] = sh[t];
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }
// blobs per hash CTA (CKZG_B200_HASH_LANES overrides, for measurements)
static int hash_lanes(uint64_t n) {
    static const int forced = getenv("CKZG_B200_HASH_LANES") ? atoi(getenv("CKZG_B200_HASH_LANES")) : 0;
    if (forced >= 1 && forced <= 32) return forced;
    (void)n;
    return 32;
}

int launch_blob_challenges(Launch& L, Fr* z, uint8_t* zy, const uint8_t* blobs, const uint8_t* commitments48, uint64_t n) {
    return launch_blob_challenges_range(L, z, zy, blobs, commitments48, n, 0, CH_BLOCKS, nullptr);
}
int blob_challenge_blocks() { return CH_BLOCKS; }
// SHA-256 blocks [k0, k1) of every blob's challenge hash; states (8 words per blob) carries the chaining value between
// launches (read if k0 > 0, written if k1 < blob_challenge_blocks(); z / zy are written by the launch that ends the hash)
int launch_blob_challenges_range(Launch& L, Fr* z, uint8_t* zy, const uint8_t* blobs, const uint8_t* commitments48, uint64_t n, int k0, int k1, uint32_t* states) {
    if (!n) return RET_OK;
    if (k0 < 0 || k1 > CH_BLOCKS || k0 >= k1 || ((k0 > 0 || k1 < CH_BLOCKS) && !states)) return RET_ERROR;
    const int lanes = hash_lanes(n);
    if (sha_fma_adds())
        blob_challenge_kernel<true><<<blocks_for(n, lanes), 64, 0, L.stream>>>(z, zy, blobs, commitments48, n, lanes, k0, k1, states);
    else
        blob_challenge_kernel<false><<<blocks_for(n, lanes), 64, 0, L.stream>>>(z, zy, blobs, commitments48, n, lanes, k0, k1, states);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "blob_challenge");
    return RET_OK;
}
int launch_z_from_bytes(Launch& L, Fr* z, uint8_t* zy, const uint8_t* z_bytes, uint64_t n, int* bad) {
    if (!n) return RET_OK;
    z_from_bytes_kernel<<<blocks_for(n, 64), 64, 0, L.stream>>>(z, zy, z_bytes, n, bad);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}
int launch_fr_from_bytes(Launch& L, Fr* out, const uint8_t* bytes32, uint64_t stride, uint64_t n, int* bad) {
    if (!n) return RET_OK;
    fr_from_bytes_kernel<<<blocks_for(n, 64), 64, 0, L.stream>>>(out, bytes32, stride, n, bad);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}
int launch_evaluate(Launch& L, Fr* y, uint8_t* zy, Fr* inv_or_null, int* m_or_null, const uint8_t* blobs, const Fr* z, uint64_t n, int* bad, int bad_stride) {
    if (!n) return RET_OK;
    static const bool use_tree = !(getenv("CKZG_B200_EVAL") && strcmp(getenv("CKZG_B200_EVAL"), "barycentric") == 0);
    if (!inv_or_null && !m_or_null && use_tree) {  // verifiers: only y is needed
        Fr* zpow = nullptr;
        KZG_CUDA_TRY(cudaMallocAsync((void**)&zpow, n * 12 * sizeof(Fr), L.stream));
        evaluate_zpow_kernel<<<blocks_for(n, 32), 32, 0, L.stream>>>(zpow, z, n);
        KZG_CUDA_TRY(cudaGetLastError());
        static const int leaves_log = (getenv("CKZG_B200_EVAL_LEAVES") && atoi(getenv("CKZG_B200_EVAL_LEAVES")) == 32) ? 5 : 4;
        if (leaves_log == 5)
            evaluate_tree_kernel<5><<<(unsigned)n, N_BLOB >> 5, 0, L.stream>>>(y, zy, blobs, zpow, L.ctx->roots_brp, bad, bad_stride);
        else
            evaluate_tree_kernel<4><<<(unsigned)n, N_BLOB >> 4, 0, L.stream>>>(y, zy, blobs, zpow, L.ctx->roots_brp, bad, bad_stride);
        KZG_CUDA_TRY(cudaGetLastError());
        KZG_CUDA_TRY(cudaFreeAsync(zpow, L.stream));
        L.count(2, "evaluate");
        return RET_OK;
    }
    Fr* total = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&total, n * sizeof(Fr), L.stream));
    evaluate_products_kernel<<<(unsigned)n, EV_THREADS, 0, L.stream>>>(total, z, L.ctx->roots_brp);
    KZG_CUDA_TRY(cudaGetLastError());
    evaluate_invert_kernel<<<blocks_for(n, 32), 32, 0, L.stream>>>(total, n);
    KZG_CUDA_TRY(cudaGetLastError());
    evaluate_kernel<<<(unsigned)n, EV_THREADS, 0, L.stream>>>(y, zy, inv_or_null, m_or_null, blobs, z, L.ctx->roots_brp, total, bad, bad_stride);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaFreeAsync(total, L.stream));
    L.count(3, "evaluate");
    return RET_OK;
}
int launch_quotient(Launch& L, uint8_t* q_scalars, const uint8_t* blobs, const Fr* z, const Fr* y, const Fr* inv, const int* m, uint64_t n) {
    if (!n) return RET_OK;
    dim3 grid(N_BLOB / 256, (unsigned)n);
    quotient_kernel<<<grid, 256, 0, L.stream>>>(q_scalars, blobs, y, inv, m);
    KZG_CUDA_TRY(cudaGetLastError());
    quotient_in_domain_kernel<<<(unsigned)n, 256, 0, L.stream>>>(q_scalars, blobs, z, y, inv, L.ctx->roots_brp, m);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "quotient");
    return RET_OK;
}
// two arrays in ONE launch: the kernel is latency bound (one 1.9k-product chain per thread, 1 warp
// per CTA), so validating commitments and proofs together costs the time of one
__global__ void g1_validate2_kernel(G1Affine* __restrict__ out_a, const uint8_t* __restrict__ in_a, G1Affine* __restrict__ out_b, const uint8_t* __restrict__ in_b, uint64_t n,
                                    uint64_t nb, int* __restrict__ bad) {
    place_record(2);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + nb) return;
    const bool second = i >= n;
    const uint64_t k = second ? i - n : i;
    const uint8_t* src = (second ? in_b : in_a) + k * 48;
    uint8_t buf[48];
    for (int q = 0; q < 48; q++) buf[q] = src[q];
    G1Affine a;
    if (!g1a_validate(a, buf)) {
        *bad = 1;
        a = g1a_inf();
    }
    (second ? out_b : out_a)[k] = a;
}
// Per-blob stage, part one, as ONE kernel: the four warps of a CTA sit on the four sub-partitions of an
// SM; warps 0/1 hash 32 blobs (challenge_warp_pair), warp 2 validates their 32 commitments and warp 3
// their 32 proofs.  All three are single-lane dependency chains of ~2 ms that saturate the issue port of
// the sub-partition they run on: as separate concurrent kernels the block scheduler stacked them on the
// same sub-partitions and each took as long as running them back to back (tools/gpu_probe.py
// "placement", profiles/r01_summary.md r01q-r01s).
template <bool FMA_ADDS>
__global__ void __launch_bounds__(128) stage1_fused_kernel(Fr* __restrict__ z_out, uint8_t* __restrict__ zy, const uint8_t* __restrict__ blobs, G1Affine* __restrict__ out_cm,
                                                           const uint8_t* __restrict__ in_cm, G1Affine* __restrict__ out_pf, const uint8_t* __restrict__ in_pf, uint64_t n,
                                                           int* __restrict__ bad, G1* __restrict__ table, int lanes) {
    __shared__ uint32_t kw[2][64][32];
    place_record(3);
    const int warp = threadIdx.x >> 5;
    if (warp < 2) {
        challenge_warp_pair<FMA_ADDS>(kw, z_out, zy, blobs, in_cm, n, lanes);
        return;
    }
    if ((int)(threadIdx.x & 31) >= lanes) return;
    const uint64_t i = (uint64_t)blockIdx.x * lanes + (threadIdx.x & 31);
    if (i >= n) return;
    const uint8_t* src = (warp == 2 ? in_cm : in_pf) + i * 48;
    uint8_t buf[48];
    for (int q = 0; q < 48; q++) buf[q] = src[q];
    G1Affine a;
    // table columns: proofs 0..n-1, commitments n..2n-1 (vmsm.cu)
    const bool ok = table ? g1a_validate_levels(a, buf, table + (warp == 2 ? n + i : i), 2 * n + 1) : g1a_validate(a, buf);
    if (!ok) {
        *bad = 1;
        a = g1a_inf();
    }
    (warp == 2 ? out_cm : out_pf)[i] = a;
}
// (A variant of this kernel with the validations on quads of lanes -- 8 more warps per CTA -- was measured in
// r02l: the extra warps share the sub-partitions of the hash pair and the stage went 2.35 -> 3.45 ms.  The
// hash is the floor of this stage; the quad validator below serves where no hash runs beside it.)
// n_a + n_b points (two byte arrays) on QUADS of lanes (g1_quad.cuh g1a_validate_levels_quad), 32 per CTA; table
// columns col_a + k / col_b + k of `npts`.  The two |z| chains of the subgroup test take a third of the time on a
// quad; more total work than one lane per point, so callers use it while the batch fits one wave.
constexpr int VQ_QUADS = 32;
__global__ void __launch_bounds__(4 * VQ_QUADS) g1_validate_levels_quad_kernel(G1Affine* __restrict__ out_a, const uint8_t* __restrict__ in_a, uint64_t na, uint64_t col_a,
                                                                               G1Affine* __restrict__ out_b, const uint8_t* __restrict__ in_b, uint64_t nb, uint64_t col_b,
                                                                               G1* __restrict__ table, uint64_t npts, int* __restrict__ bad) {
    extern __shared__ __align__(16) unsigned char vq_smem[];
    const uint64_t g = (uint64_t)blockIdx.x * VQ_QUADS + (threadIdx.x >> 2);
    if (g >= na + nb) return;
    const bool second = g >= na;
    const uint64_t k = second ? g - na : g;
    QuadValidate* W = reinterpret_cast<QuadValidate*>(vq_smem) + (threadIdx.x >> 2);
    G1Affine* out = second ? out_b : out_a;
    G1Affine* oa = out ? out + k : nullptr;
    const bool ok = g1a_validate_levels_quad(oa, (second ? in_b : in_a) + k * 48, table + (second ? col_b : col_a) + k, npts, W);
    if (!ok && (threadIdx.x & 3) == 0) {
        *bad = 1;
        if (oa) *oa = g1a_inf();
    }
}
int launch_g1_validate_levels_ab(Launch& L, G1Affine* out_a, const uint8_t* in_a, uint64_t na, uint64_t col_a, G1Affine* out_b, const uint8_t* in_b, uint64_t nb, uint64_t col_b,
                                 G1* table, uint64_t npts, int* bad) {
    if (!(na + nb)) return RET_OK;
    g1_validate_levels_quad_kernel<<<blocks_for(na + nb, VQ_QUADS), 4 * VQ_QUADS, VQ_QUADS * sizeof(QuadValidate), L.stream>>>(out_a, in_a, na, col_a, out_b, in_b, nb, col_b, table,
                                                                                                                           npts, bad);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "g1_validate");
    return RET_OK;
}

int launch_stage1_fused(Launch& L, Fr* z, uint8_t* zy, const uint8_t* blobs, G1Affine* out_cm, const uint8_t* in_cm, G1Affine* out_pf, const uint8_t* in_pf, uint64_t n, int* bad,
                        G1* table) {
    if (!n) return RET_OK;
    const int lanes = hash_lanes(n);
    if (sha_fma_adds())
        stage1_fused_kernel<true><<<blocks_for(n, lanes), 128, 0, L.stream>>>(z, zy, blobs, out_cm, in_cm, out_pf, in_pf, n, bad, table, lanes);
    else
        stage1_fused_kernel<false><<<blocks_for(n, lanes), 128, 0, L.stream>>>(z, zy, blobs, out_cm, in_cm, out_pf, in_pf, n, bad, table, lanes);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "hash+validate");
    return RET_OK;
}
// validation of n commitments (columns n..2n-1) and n proofs (columns 0..n-1) that also writes their table columns
__global__ void g1_validate2_levels_kernel(G1Affine* __restrict__ out_cm, const uint8_t* __restrict__ in_cm, G1Affine* __restrict__ out_pf, const uint8_t* __restrict__ in_pf,
                                           uint64_t n, int* __restrict__ bad, G1* __restrict__ table) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    time_record(2u);
    const bool is_cm = i >= n;
    const uint64_t k = is_cm ? i - n : i;
    const uint8_t* src = (is_cm ? in_cm : in_pf) + k * 48;
    uint8_t buf[48];
    for (int q = 0; q < 48; q++) buf[q] = src[q];
    G1Affine a;
    if (!g1a_validate_levels(a, buf, table + i, 2 * n + 1)) {
        *bad = 1;
        a = g1a_inf();
    }
    (is_cm ? out_cm : out_pf)[k] = a;
}
int launch_g1_validate2_levels(Launch& L, G1Affine* out_cm, const uint8_t* in_cm, G1Affine* out_pf, const uint8_t* in_pf, uint64_t n, int* bad, G1* table) {
    if (!n) return RET_OK;
    g1_validate2_levels_kernel<<<blocks_for(2 * n, 32), 32, 0, L.stream>>>(out_cm, in_cm, out_pf, in_pf, n, bad, table);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "g1_validate");
    return RET_OK;
}
int debug_set_timer_buffer(uint32_t* dev_buf) {
    KZG_CUDA_TRY(cudaMemcpyToSymbol(g_time_buf, &dev_buf, sizeof(dev_buf)));
    return RET_OK;
}
int debug_set_placement_buffer(uint32_t* dev_buf) {
    KZG_CUDA_TRY(cudaMemcpyToSymbol(g_place_buf, &dev_buf, sizeof(dev_buf)));
    return RET_OK;
}

int launch_g1_validate2(Launch& L, G1Affine* out_a, const uint8_t* in_a, G1Affine* out_b, const uint8_t* in_b, uint64_t n, int* bad) {
    return launch_g1_validate_ab(L, out_a, in_a, n, out_b, in_b, n, bad);
}
int launch_g1_validate_ab(Launch& L, G1Affine* out_a, const uint8_t* in_a, uint64_t n, G1Affine* out_b, const uint8_t* in_b, uint64_t nb, int* bad) {
    if (!(n + nb)) return RET_OK;
    g1_validate2_kernel<<<blocks_for(n + nb, 32), 32, 0, L.stream>>>(out_a, in_a, out_b, in_b, n, nb, bad);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "g1_validate");
    return RET_OK;
}

int launch_g1_validate(Launch& L, G1Affine* out, const uint8_t* bytes48, uint64_t n, int* bad, int bad_stride) {
    if (!n) return RET_OK;
    g1_validate_kernel<<<blocks_for(n, 32), 32, 0, L.stream>>>(out, bytes48, n, bad, bad_stride);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "g1_validate");
    return RET_OK;
}
__global__ void fr_to_bytes_kernel(uint8_t* __restrict__ out, const Fr* __restrict__ in, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_fr_be(out + 32 * i, in[i]);
}
int launch_fr_to_bytes(Launch& L, uint8_t* out32, const Fr* in, uint64_t n) {
    if (!n) return RET_OK;
    fr_to_bytes_kernel<<<blocks_for(n, 64), 64, 0, L.stream>>>(out32, in, n);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "fr_to_bytes");
    return RET_OK;
}

int launch_r_from_digest(Launch& L, Fr* r, const uint8_t* digest32) {
    r_from_digest_kernel<<<1, 32, 0, L.stream>>>(r, digest32);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "r_from_digest");
    return RET_OK;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t rlc_fold(uint64_t n) { return (2 * n) / FOLD_IN + 2; }
size_t rlc_scratch_bytes(uint64_t n) {
    size_t fold = rlc_fold(n);
    return al256(n * 32) * 2 + al256(n * sizeof(Fr)) + al256(32) + al256(sizeof(G1)) + al256((n + fold) * sizeof(G1)) +
           al256((2 * n + fold + 2) * sizeof(G1));
}

__global__ void lift_affine_kernel(G1* out, const G1Affine* in, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = g1_from_affine(in[i]);
}
int launch_lift_affine(Launch& L, G1* out, const G1Affine* in, uint64_t n) {
    if (!n) return RET_OK;
    lift_affine_kernel<<<blocks_for(n, 64), 64, 0, L.stream>>>(out, in, n);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}

int launch_g1_sum(Launch& L, G1* out, G1* in, uint64_t n) {
    // ping-pong inside `in` is not possible (block outputs overlap later inputs only after they are
    // consumed by the same block), so fold in place level by level: block b writes slot b, which was
    // read by block 0 of the same launch only if b < SUM_THREADS*SUM_PER -- use a second buffer.
    // Callers pass scratch of at least ceil(n / 1024) + 1 points after `in`'s n points.
    G1* a = in;
    G1* b = in + n;
    uint64_t m = n;
    while (m > 1) {
        unsigned blocks = blocks_for(m, SUM_THREADS * SUM_PER);
        g1_sum_kernel<<<blocks, SUM_THREADS, 0, L.stream>>>(b, a, m);
        KZG_CUDA_TRY(cudaGetLastError());
        L.count(1, "g1_sum");
        G1* t = a;
        a = b;
        b = t;
        m = blocks;
    }
    KZG_CUDA_TRY(cudaMemcpyAsync(out, a, sizeof(G1), cudaMemcpyDeviceToDevice, L.stream));
    return RET_OK;
}

int launch_rlc(Launch& L, G1* out2, const G1Affine* commitments, const G1Affine* proofs, const Fr* z, const Fr* y, const Fr* r, bool use_r, uint64_t first,
               uint64_t n, void* scratch) {
    uint8_t* ws = (uint8_t*)scratch;
    uint32_t* s1 = (uint32_t*)ws; ws += al256(n * 32);
    uint32_t* s2 = (uint32_t*)ws; ws += al256(n * 32);
    Fr* ty = (Fr*)ws; ws += al256(n * sizeof(Fr));
    uint32_t* ysum = (uint32_t*)ws; ws += al256(32);
    G1* Y = (G1*)ws; ws += al256(sizeof(G1));
    // U: n points + fold scratch ; T: 2n points + fold scratch
    size_t fold = rlc_fold(n);
    G1* U = (G1*)ws; ws += al256((n + fold) * sizeof(G1));
    G1* T = (G1*)ws;

    rlc_scalars_kernel<<<blocks_for(n, 64), 64, 0, L.stream>>>(s1, s2, ty, z, y, r, use_r, first, n);
    KZG_CUDA_TRY(cudaGetLastError());
    fr_sum_kernel<<<1, 256, 0, L.stream>>>(ysum, ty, n);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "rlc_scalars");
    rlc_points_kernel<<<blocks_for(3 * n + 1, 64), 64, 0, L.stream>>>(U, T, Y, commitments, proofs, s1, s2, ysum, n);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "rlc_points");
    G1* PU = U + n;
    G1* PT = T + 2 * n;
    const unsigned fb = blocks_for(2 * n, FOLD_IN);
    rlc_fold_kernel<<<dim3(fb, 2), FOLD_THREADS, 0, L.stream>>>(PU, PT, U, T, n);
    KZG_CUDA_TRY(cudaGetLastError());
    rlc_final_kernel<<<1, 64, 0, L.stream>>>(out2, PU, blocks_for(n, FOLD_IN), PT, fb, Y);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(2, "g1_sum");
    return RET_OK;
}

}  // namespace kzg
