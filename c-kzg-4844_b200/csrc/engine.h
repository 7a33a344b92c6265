// Internal (C++) interface of the CUDA engine: the device-resident context and the kernel launchers
// that the C-ABI layer (api.cu) strings together.  Not installed; the public surface is include/*.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "g1.cuh"
#include "hostpool.h"

namespace kzg {

// c-kzg return codes (src/common/ret.h:24-29)
enum { RET_OK = 0, RET_BADARGS = 1, RET_ERROR = 2, RET_MALLOC = 3 };

constexpr int N_BLOB = 4096;        // FIELD_ELEMENTS_PER_BLOB (src/eip4844/blob.h:29)
constexpr int N_EXT = 8192;         // FIELD_ELEMENTS_PER_EXT_BLOB (blob.h:42)
constexpr int BLOB_BYTES = 131072;  // BYTES_PER_BLOB

// ---- fixed-base MSM geometry (DESIGN.md "MSM") -------------------------------------------------
// Signed c-bit digits; the table holds 2^(c*j) * G_i for every window j, so all windows of a blob
// share one set of 2^(c-1) buckets and no doublings are needed.
constexpr int MSM_C = 11;
constexpr int MSM_W = 24;                    // ceil(256 / 11): covers the carry out of bit 254
constexpr int MSM_NB = 1 << (MSM_C - 1);     // 1024 buckets per blob
constexpr int MSM_ENTRIES = MSM_W * N_BLOB;  // 98304 (digit, point) pairs per blob

struct PairingLines;  // pairing.cuh

struct Prof {
    std::mutex mu;
    int level = 0;  // 0 off, 1 whole-call events only, 2 per-kernel events (concurrent stages run serially)
    static constexpr int MAXK = 64;
    const char* names[MAXK];
    double ms[MAXK];
    uint64_t cnt[MAXK];
    int nk = 0;
    double call_ms = 0;  // begin-to-end device time of all calls
    uint64_t calls = 0;
    void add(const char* name, double v) {
        for (int i = 0; i < nk; i++)
            if (names[i] == name || strcmp(names[i], name) == 0) {
                ms[i] += v;
                cnt[i]++;
                return;
            }
        if (nk < MAXK) {
            names[nk] = name;
            ms[nk] = v;
            cnt[nk] = 1;
            nk++;
        }
    }
};

struct Ctx {
    int device = 0;
    std::atomic<uint64_t> launches{0};
    Prof prof;

    // trusted setup, device resident (Montgomery form)
    G1Affine* g1_monomial = nullptr;      // [4096]
    G1Affine* g1_lagrange_brp = nullptr;  // [4096] bit-reversed Lagrange points
    G1Affine* msm_table = nullptr;        // [MSM_W][4096]: 2^(c*j) * g1_lagrange_brp[i]
    Fr* roots_brp = nullptr;              // [8192] brp(roots_of_unity[0..8191]) (first 4096 = blob domain)
    Fr* roots = nullptr;                  // [8193] w^i
    void* g2_lines = nullptr;             // precomputed Miller-loop lines for G2 gen, [tau]G2, [tau^64]G2
    void* pairing_tables = nullptr;       // lane schedules of the cooperative pairing arithmetic (pairing.cu pairing_tables_kernel)
    void* g2_points = nullptr;            // [65] affine G2 (Fp2 coordinates)
    void* rec_shiftA = nullptr;           // 7^k / 8192   (recover.cu)
    void* rec_shiftB = nullptr;           // 7^-k / 8192
    void* fk_table = nullptr;             // FK20 fixed-base multiples, cells.h (3.2 - 35 GB)
    int fk_c = 8;                         // window width of fk_table
    void* commit_table = nullptr;         // direct (bucket-free) multiples of the Lagrange points, msm_direct.cu (18 - 61 GB)
    int commit_c = 0;                     // its window width; 0 = not built, the bucket MSM (msm_table) serves
    std::once_flag commit_once;           // both tables are built on first use (api.cu plan_*_window)
    std::mutex fk_mu;
    std::atomic<bool> fk_ready{false};
    G1* g_levels = nullptr;               // [18] table levels of -G1 generator (vmsm.cu)
    G1* mono_levels = nullptr;            // [18][64] table levels of -[tau^j]G1, j < 64 (verify_cells.cu)
    uint64_t precompute = 0;
    void* coalescer = nullptr;            // call-coalescing front end (coalesce.cu)

    // In-library multi-device (multi.cu): with CKZG_B200_DEVICES=0,1,... the context returned to the caller is the
    // PRIMARY (peers[0] == this) and owns one ordinary context per further device; the batched entry points shard
    // HOST-memory batches over them.  Empty for a single-device context and for the peers themselves.
    std::vector<Ctx*> peers;
    // host worker threads (parallel memcpy into pinned staging, per-device / per-sub-batch fan-out); created on demand
    std::mutex pool_mu;
    HostPool* pool = nullptr;
    HostPool* host_pool() {
        std::lock_guard<std::mutex> g(pool_mu);
        if (!pool) pool = new HostPool(host_threads_default() - 1);  // the caller is the last worker
        return pool;
    }

    // Small pool of pinned host buffers for the device->host hops inside a call (a pageable destination
    // makes cudaMemcpyAsync stage through the driver and block).  Buffers are reused across calls and
    // released with the context.
    // Streams of finished calls, kept for the next one.  A host-pointer batch verification uses nine streams; creating
    // them delayed the first copy by 0.3 ms and destroying them kept the caller 0.5 ms after the verdict had arrived
    // (host marks, profiles/e2e_probe_R3c.log: 14.47 ms per call of which 13.88 between the device events).  A call's
    // streams are idle when it returns them (its destructor synchronises), so reuse needs no further ordering.
    // kind 0: blocking (call streams), 1: non-blocking (side streams).  CKZG_B200_STREAM_POOL=0: create / destroy per call.
    std::mutex stream_mu;
    std::vector<cudaStream_t> stream_free[2];
    static bool stream_pool_on() {
        static const bool on = !(getenv("CKZG_B200_STREAM_POOL") && atoi(getenv("CKZG_B200_STREAM_POOL")) == 0);
        return on;
    }
    cudaStream_t stream_acquire(int kind, bool pooled = true) {
        if (pooled && stream_pool_on()) {
            std::lock_guard<std::mutex> g(stream_mu);
            if (!stream_free[kind].empty()) {
                cudaStream_t s = stream_free[kind].back();
                stream_free[kind].pop_back();
                return s;
            }
        }
        cudaStream_t s = nullptr;
        if (cudaStreamCreateWithFlags(&s, kind ? cudaStreamNonBlocking : cudaStreamDefault) != cudaSuccess) return nullptr;
        return s;
    }
    void stream_release(int kind, cudaStream_t s, bool pooled = true) {
        if (pooled && stream_pool_on()) {
            std::lock_guard<std::mutex> g(stream_mu);
            if (stream_free[kind].size() < 64) {
                stream_free[kind].push_back(s);
                return;
            }
        }
        cudaStreamDestroy(s);
    }
    void stream_pool_destroy() {
        std::lock_guard<std::mutex> g(stream_mu);
        for (int k = 0; k < 2; k++) {
            for (cudaStream_t s : stream_free[k]) cudaStreamDestroy(s);
            stream_free[k].clear();
        }
    }
    std::mutex pin_mu;
    std::vector<std::pair<void*, size_t>> pin_free;
    void* pin_acquire(size_t bytes, size_t* got) {
        {
            std::lock_guard<std::mutex> g(pin_mu);
            // best fit: a 48-byte request must not walk off with the 16 MB staging block of a coalesced batch
            size_t best = pin_free.size();
            for (size_t i = 0; i < pin_free.size(); i++)
                if (pin_free[i].second >= bytes && (best == pin_free.size() || pin_free[i].second < pin_free[best].second)) best = i;
            if (best != pin_free.size()) {
                void* p = pin_free[best].first;
                *got = pin_free[best].second;
                pin_free.erase(pin_free.begin() + best);
                return p;
            }
        }
        void* p = nullptr;
        size_t cap = bytes < 4096 ? 4096 : bytes;
        // portable: the multi-device paths hand one staging block to copies issued on several devices
        if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) return nullptr;
        *got = cap;
        return p;
    }
    void pin_release(void* p, size_t cap) {
        std::lock_guard<std::mutex> g(pin_mu);
        if (pin_free.size() < 16) {
            pin_free.push_back({p, cap});
            return;
        }
        // keep the pool small: drop the smallest
        size_t k = 0;
        for (size_t i = 1; i < pin_free.size(); i++)
            if (pin_free[i].second < pin_free[k].second) k = i;
        if (pin_free[k].second < cap) {
            cudaFreeHost(pin_free[k].first);
            pin_free[k] = {p, cap};
        } else {
            cudaFreeHost(p);
        }
    }
};

// Optional per-kernel timing (ckzg_b200_profile_*): when enabled, every count() drops a CUDA event on
// the call's stream; the time between consecutive events is attributed to the kernel(s) just
// launched.  Events sit on the launching stream, so this is what bench.py uses for the roofline.
struct ProfTrace {
    std::vector<std::pair<cudaEvent_t, const char*>> ev;
};

struct Launch {
    Ctx* ctx;
    cudaStream_t stream;
    ProfTrace* trace;
    void count(int n = 1, const char* name = nullptr) {
        ctx->launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
        if (trace) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) == cudaSuccess) {
                cudaEventRecord(e, stream);
                trace->ev.push_back({e, name ? name : "other"});
            }
        }
    }
};

#define KZG_CUDA_TRY(expr)                         \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) {                   \
            kzg::note_cuda_error(_e, __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? kzg::RET_MALLOC : kzg::RET_ERROR; \
        }                                          \
    } while (0)

void note_cuda_error(cudaError_t e, const char* file, int line);

// cudaFuncSetAttribute acts on the CURRENT device only: a process that drives several devices (multi.cu, or several
// contexts) must opt in on each of them, so the "done once" flag is one bit per device.
#define KZG_FUNC_ATTR_PER_DEVICE(kernel, attr, value)                    \
    do {                                                                 \
        static std::atomic<uint64_t> _done{0};                           \
        int _dev = 0;                                                    \
        cudaGetDevice(&_dev);                                            \
        const uint64_t _bit = 1ull << (_dev & 63);                       \
        if (!(_done.load(std::memory_order_acquire) & _bit)) {           \
            KZG_CUDA_TRY(cudaFuncSetAttribute(kernel, attr, value));     \
            _done.fetch_or(_bit, std::memory_order_release);             \
        }                                                                \
    } while (0)

// ---- setup.cu ----------------------------------------------------------------------------------
// out[i] = uncompress(bytes[48*perm(i)]) ; ok_flag (device int) is cleared on any failure.
int launch_g1_uncompress(Launch& L, G1Affine* out, const uint8_t* bytes, int n, bool bit_reverse, bool check_subgroup, int* d_bad);
int launch_msm_table(Launch& L, G1Affine* table, const G1Affine* base);
int launch_roots(Launch& L, Fr* roots, Fr* roots_brp);

// ---- msm.cu ------------------------------------------------------------------------------------
// scalars: n x 4096 x 32 bytes; `big_endian_bytes` = wire blobs (canonical check -> bad[blob] = 1),
// else plain little-endian limbs (already reduced).  result: n XYZZ points (device).
struct MsmWorkspace {
    uint32_t* entries = nullptr;  // [n][MSM_ENTRIES]
    uint32_t* starts = nullptr;   // [n][MSM_NB + 1]
    G1* buckets = nullptr;        // [n][parts][MSM_NB]
    int parts = 1;
    uint64_t n = 0;
};
size_t msm_workspace_bytes(uint64_t n, int parts);
int msm_pick_parts(uint64_t n);
int launch_msm(Launch& L, G1* result, const uint8_t* scalars, bool big_endian_bytes, uint64_t n, const G1Affine* table, int* d_bad, void* workspace, int parts);
// ---- msm_direct.cu: the same sums from the direct table (Ctx::commit_table), no sort / buckets ----------
// builds Ctx::commit_table on first use if the free HBM allows (else leaves it null: bucket form)
int msm_direct_ensure(Ctx* c);
int plan_commit_window();
int plan_fk_window();
size_t msm_direct_workspace_bytes(uint64_t n);
int launch_msm_direct(Launch& L, G1* result, const uint8_t* scalars, bool big_endian_bytes, uint64_t n, int* d_bad, void* workspace);
// ---- msm_affine.cu: the same table sums added pairwise in affine coordinates with batched inversions (batches) ----
size_t msm_affine_workspace_bytes(uint64_t n, int c);
int launch_msm_affine(Launch& L, G1* result, const uint8_t* scalars, bool big_endian_bytes, uint64_t n, int* d_bad, void* workspace);
// points -> canonical 48-byte compression (one thread per point)
int launch_g1_compress(Launch& L, uint8_t* out48, const G1* pts, uint64_t n);

// ---- pairing.cu: measurement hook ----
int debug_pairing_probe(Ctx* c, long long* ticks_host, int* ok_host, const uint8_t* two48_host, int reps);
// ---- selftest.cu ---------------------------------------------------------------------------------
int selftest_field(int op, uint32_t* out, const uint32_t* a, const uint32_t* b, uint64_t n);
int selftest_mulbench(int ilp, int iters, int blocks, int threads, float* ms_out);
int selftest_g1(int op, uint8_t* out48, int* ok_out, const uint8_t* p48, const uint32_t* k, const uint8_t* q48, uint64_t n);

}  // namespace kzg
