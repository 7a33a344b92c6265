// Shared pieces of the variable-base bucket MSM (vmsm.cu): job descriptors, digit storage, launchers.
// Used by the blob verifier (vmsm.cu launch_rlc_vmsm) and the cell verifier (verify_cells.cu).
#pragma once
#include "g1.cuh"
#include "verify.h"

namespace kzg {

constexpr int VC = 8;                    // digit width (one byte of a 64-bit base-|z| digit)
constexpr int VW = 9;                    // table levels per base (the balanced digits use 8 of them)
static_assert(VMSM_LEVELS == 2 * VW && VMSM_LEVELS == G1_LEVELS, "table layout: 9 levels of P, 9 levels of [|z|]P");
constexpr int VNB = 1 << (VC - 1);       // 128 buckets (signed digits, magnitude 1..128)
constexpr int VSORT_THREADS = 256;
constexpr int VSORT_WARPS = VSORT_THREADS / 32;
constexpr int VSORT_CTAS = 64;           // CTAs per MSM in the counting sort
constexpr uint32_t VCAP = 8;             // list entries folded by one accumulate thread
constexpr int VACC_THREADS = 128;
constexpr int VCOMB_THREADS = 128;

struct VmsmJob {
    const uint32_t* halves;  // [nh][2] 64-bit base-|z| digits, index h = 4 * point + quarter
    uint32_t nh;
    uint32_t max_items;
    uint32_t* entries;       // [nh * VW]
    uint32_t* starts;        // [VNB + 1]
    uint32_t* item_start;    // [VNB + 1]
    uint32_t* item_bucket;   // [max_items]
    G1* partial;             // [max_items]
    G1* combined;            // [VNB]
    uint32_t* ctahist;       // [VSORT_CTAS][VNB]
    G1* scan_tmp;            // [VNB]
};
struct VmsmJobs {
    VmsmJob j[2];
};

// One point's scalar -> four balanced base-|z| digits (magnitude | sign in bit 63), 32 bytes at index `point`.
__device__ __forceinline__ void store_halves(uint32_t* hB, size_t point, const uint32_t k[8]) {
    int64_t sd[4];
    basez_split(sd, k);
    uint64_t a[4];  // magnitude (< 2^63) | sign in bit 63
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = sd[i] < 0 ? ((uint64_t)(-sd[i]) | (1ull << 63)) : (uint64_t)sd[i];
    uint4* dst = reinterpret_cast<uint4*>(hB) + 2 * point;
    dst[0] = make_uint4((uint32_t)a[0], (uint32_t)(a[0] >> 32), (uint32_t)a[1], (uint32_t)(a[1] >> 32));
    dst[1] = make_uint4((uint32_t)a[2], (uint32_t)(a[2] >> 32), (uint32_t)a[3], (uint32_t)(a[3] >> 32));
}

struct Digest8 {
    uint32_t h[8];  // big-endian words of the SHA-256 digest
};
// r = hash_to_bls_field(digest) (src/common/bytes.c:123): the digest is below 2^256 < 3r
__device__ __forceinline__ Fr fr_from_digest_words(const uint32_t h[8]) {
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
// digest bytes (host) -> kernel-argument form
inline Digest8 digest8_from_bytes(const uint8_t* d) {
    Digest8 dg;
    for (int i = 0; i < 8; i++) dg.h[i] = ((uint32_t)d[4 * i] << 24) | ((uint32_t)d[4 * i + 1] << 16) | ((uint32_t)d[4 * i + 2] << 8) | (uint32_t)d[4 * i + 3];
    return dg;
}

// ---- host side (vmsm.cu) ---------------------------------------------------------------------------
size_t vmsm_job_bytes(uint64_t nh);
// carves one job's arrays out of `ws` (vmsm_job_bytes(nh) bytes); returns the advanced pointer
uint8_t* vmsm_job_carve(VmsmJob& J, uint8_t* ws, const uint32_t* halves, uint64_t nh);
// sort / accumulate / combine / reduce for the two jobs over one table of `npts` columns:
// out2[k] = sum over job k's digits.  Job 1 must be the larger one (it sizes the accumulate grid).
int launch_vmsm_jobs(Launch& L, G1* out2, const VmsmJobs& jobs, const G1* table, uint32_t npts);
// levels[j * stride + i], j < 18: the table levels of (+-)pts[i] (setup-time tables of fixed bases)
int launch_vmsm_point_levels(Launch& L, G1* levels, const G1Affine* pts, uint32_t n, bool negate);

}  // namespace kzg
