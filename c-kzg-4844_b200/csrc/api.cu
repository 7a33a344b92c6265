// The C ABI of the engine (include/ckzg_b200.h): context lifecycle and the batched entry points.
// Host-side orchestration only -- every field/curve/hash operation runs in a CUDA kernel; there is
// no CPU arithmetic path in this library (a missing/failed device surfaces as C_KZG_ERROR).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "../../include/ckzg_b200.h"
#include "engine.h"
#include "verify.h"
#include "cells.h"
#include "call.h"

namespace kzg {

// Hardware work queues.  A process gets CUDA_DEVICE_MAX_CONNECTIONS (default 8) hardware queues per device and its
// streams are spread over them.  A host-pointer verification uses a dozen streams (copies, four to eight hash streams,
// copy-back, the call stream), and with eight or more queues active the host interface came back to queues late:
// events recorded behind finished kernels fired up to 11 ms after the kernels had ended, erratically -- one arrangement
// fast, its neighbour 30 % slower, two concurrent callers anywhere between 95 k and 395 k blobs/s (profiles/
// e2e_probe_R2k ... R2y.log).  With FOUR queues every arrangement measured was fast and stable (R2y: 14.9 ms per 4096-blob
// call, two callers 364 k blobs/s), with 32 every one was slow.  The variable only counts if it is set before the CUDA
// context exists, so the library sets it (without overriding the user's value) when it is loaded; a host process that
// initialises CUDA before loading the library (PyTorch) sets it itself (bench.py, tests/conftest.py).
__attribute__((constructor)) static void ckzg_b200_default_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "4", 0); }

cudaStream_t& caller_stream_tls() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}

void note_cuda_error(cudaError_t e, const char* file, int line) {
    if (getenv("CKZG_B200_DEBUG")) fprintf(stderr, "[ckzg_b200] CUDA error %d (%s) at %s:%d\n", (int)e, cudaGetErrorString(e), file, line);
}

// The two big fixed-base tables are built on FIRST USE of the API that needs them (a verifier never pays for
// them) and sized from the HBM that is free at that moment -- several contexts can share a device:
//   commitment table (msm_direct.cu): c = 14 -> 61 GB | 13 -> 32 GB | 12 -> 18 GB | none (bucket MSM, msm.cu)
//   FK20 table (fk20.cu):             c = 12 -> 35 GB | 10 -> 10.5 GB | 8 -> 3.2 GB
// CKZG_B200_COMMIT_WINDOW (0 = none, 10..14) / CKZG_B200_FK_WINDOW (8, 10, 12) pin the choice.
static double free_gb() {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
    return (double)free_b / (double)(1ull << 30);
}
int plan_commit_window() {
    if (const char* env = getenv("CKZG_B200_COMMIT_WINDOW")) {
        const int v = atoi(env);
        if (v == 0 || (v >= 10 && v <= 14)) return v;
    }
    const double gb = free_gb();
    return gb >= 150 ? 14 : gb >= 90 ? 13 : gb >= 60 ? 12 : 0;
}
int plan_fk_window() {
    if (const char* env = getenv("CKZG_B200_FK_WINDOW")) {
        const int v = atoi(env);
        if (v == 8 || v == 10 || v == 12) return v;
    }
    const double gb = free_gb();
    return gb >= 70 ? 12 : gb >= 30 ? 10 : 8;
}

static int ctx_build(Ctx* c, const uint8_t* g1_mono, const uint8_t* g1_lag, const uint8_t* g2_mono) {
    Call call(c);
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();

    KZG_CUDA_TRY(cudaMalloc((void**)&c->g1_monomial, N_BLOB * sizeof(G1Affine)));
    KZG_CUDA_TRY(cudaMalloc((void**)&c->g1_lagrange_brp, N_BLOB * sizeof(G1Affine)));
    KZG_CUDA_TRY(cudaMalloc((void**)&c->msm_table, (size_t)MSM_W * N_BLOB * sizeof(G1Affine)));
    KZG_CUDA_TRY(cudaMalloc((void**)&c->roots, (N_EXT + 1) * sizeof(Fr)));
    KZG_CUDA_TRY(cudaMalloc((void**)&c->roots_brp, N_EXT * sizeof(Fr)));

    const uint8_t *d_mono, *d_lag;
    int* d_bad;
    TRY(call.stage_in(&d_mono, g1_mono, 48 * N_BLOB, CKZG_B200_HOST));
    TRY(call.stage_in(&d_lag, g1_lag, 48 * N_BLOB, CKZG_B200_HOST));
    TRY(call.alloc(&d_bad, 1));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), call.stream));

    // setup.c:447-466: decompress without subgroup check; setup.c:488: brp of the Lagrange points
    TRY(launch_g1_uncompress(L, c->g1_monomial, d_mono, N_BLOB, false, false, d_bad));
    TRY(launch_g1_uncompress(L, c->g1_lagrange_brp, d_lag, N_BLOB, true, false, d_bad));
    TRY(launch_roots(L, c->roots, c->roots_brp));
    TRY(launch_msm_table(L, c->msm_table, c->g1_lagrange_brp));
    TRY(setup_g2_and_lines(call.stream, L, c, g2_mono, d_bad));
    KZG_CUDA_TRY(cudaMalloc((void**)&c->g_levels, VMSM_LEVELS * sizeof(G1)));
    TRY(launch_vmsm_generator_levels(L, c->g_levels));
    TRY(setup_verify_cells(L, c));
    TRY(recover_setup(L, c));

    int bad = 0;
    KZG_CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    if (bad) return RET_BADARGS;
    // setup.c:339-358 is_trusted_setup_in_lagrange_form: e(L[1], G2[0]) == e(L[0], G2[1]) => monomial
    int is_monomial = 0;
    TRY(setup_is_monomial_form(call.stream, L, c, g1_lag, &is_monomial));
    if (is_monomial) return RET_BADARGS;
    return RET_OK;
}

static void ctx_free(Ctx* c) {
    if (!c) return;
    for (size_t d = 1; d < c->peers.size(); d++) ctx_free(c->peers[d]);  // peers[0] is the context itself
    c->peers.clear();
    coalescer_destroy(c);
    delete c->pool;
    c->pool = nullptr;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    c->stream_pool_destroy();
    cudaFree(c->g1_monomial);
    cudaFree(c->g1_lagrange_brp);
    cudaFree(c->msm_table);
    cudaFree(c->roots);
    cudaFree(c->roots_brp);
    cudaFree(c->g2_lines);
    cudaFree(c->pairing_tables);
    cudaFree(c->g2_points);
    cudaFree(c->fk_table);
    cudaFree(c->commit_table);
    cudaFree(c->g_levels);
    cudaFree(c->mono_levels);
    cudaFree(c->rec_shiftA);
    cudaFree(c->rec_shiftB);
    for (auto& b : c->pin_free) cudaFreeHost(b.first);
    c->pin_free.clear();
    if (prev >= 0) cudaSetDevice(prev);
    delete c;
}

}  // namespace kzg

using namespace kzg;

struct ckzg_b200_ctx {
    Ctx c;
};

extern "C" {

// one ordinary context on one device
static int ctx_create_single(Ctx** out, const uint8_t* g1_monomial_bytes, const uint8_t* g1_lagrange_bytes, const uint8_t* g2_monomial_bytes, uint64_t precompute, int device) {
    Ctx* c = new (std::nothrow) Ctx();
    if (!c) return RET_MALLOC;
    {
        // keep the stream-ordered pool's memory cached between calls (default threshold 0 hands every
        // workspace back to the driver at each sync: measured 2x on batched commitments)
        int prev = -1;
        cudaGetDevice(&prev);
        cudaMemPool_t pool;
        if (cudaSetDevice(device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        if (prev >= 0) cudaSetDevice(prev);
    }
    c->device = device;
    c->precompute = precompute;
    int rc = ctx_build(c, g1_monomial_bytes, g1_lagrange_bytes, g2_monomial_bytes);
    if (rc) {
        ctx_free(c);
        return rc;
    }
    coalescer_create(c);
    *out = c;
    return RET_OK;
}

// CKZG_B200_DEVICES = "all" | "0,1,2,3": the devices one context spans (multi.cu).  Empty / unset: one device.
static std::vector<int> devices_from_env(int ndev) {
    std::vector<int> v;
    const char* env = getenv("CKZG_B200_DEVICES");
    if (!env || !*env) return v;
    if (strcmp(env, "all") == 0) {
        for (int d = 0; d < ndev; d++) v.push_back(d);
        return v;
    }
    for (const char* p = env; *p;) {
        char* end = nullptr;
        long d = strtol(p, &end, 10);
        if (end == p) break;
        // a device may be listed twice (two replicas on one GPU): the one-GPU way to exercise the multi-device paths
        if (d >= 0 && d < ndev && v.size() < 16) v.push_back((int)d);
        p = (*end == ',') ? end + 1 : end;
        if (*end && *end != ',') break;
    }
    return v;
}

int ckzg_b200_ctx_create(ckzg_b200_ctx** out, const uint8_t* g1_monomial_bytes, const uint8_t* g1_lagrange_bytes, const uint8_t* g2_monomial_bytes, uint64_t precompute, int device) {
    if (!out || !g1_monomial_bytes || !g1_lagrange_bytes || !g2_monomial_bytes) return RET_BADARGS;
    *out = nullptr;
    if (precompute > 15) return RET_BADARGS;  // setup.c:411
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        note_cuda_error(cudaErrorNoDevice, __FILE__, __LINE__);
        return RET_ERROR;  // no CPU fallback, by design
    }
    std::vector<int> devs;
    if (device < 0) {
        devs = devices_from_env(ndev);
        const char* env = getenv("CKZG_B200_DEVICE");
        if (!devs.empty())
            device = devs[0];
        else if (env)
            device = atoi(env);
        else if (cudaGetDevice(&device) != cudaSuccess)
            return RET_ERROR;
    }
    if (device >= ndev || device < 0) return RET_BADARGS;
    Ctx* c = nullptr;
    int rc = ctx_create_single(&c, g1_monomial_bytes, g1_lagrange_bytes, g2_monomial_bytes, precompute, device);
    if (rc) return rc;
    if (devs.size() > 1) {
        // one ordinary context per further device, built concurrently (each decompresses the setup on its own GPU)
        std::vector<Ctx*> peers(devs.size(), nullptr);
        std::vector<int> rcs(devs.size(), RET_OK);
        std::vector<std::thread> th;
        peers[0] = c;
        for (size_t d = 1; d < devs.size(); d++)
            th.emplace_back([&, d] { rcs[d] = ctx_create_single(&peers[d], g1_monomial_bytes, g1_lagrange_bytes, g2_monomial_bytes, precompute, devs[d]); });
        for (auto& t : th) t.join();
        for (size_t d = 1; d < devs.size() && !rc; d++) rc = rcs[d];
        if (rc) {
            for (size_t d = 1; d < devs.size(); d++)
                if (peers[d]) ctx_free(peers[d]);
            ctx_free(c);
            return rc;
        }
        c->peers = peers;
        if (getenv("CKZG_B200_DEBUG")) {
            fprintf(stderr, "[ckzg_b200] context spans %zu devices:", devs.size());
            for (int d : devs) fprintf(stderr, " %d", d);
            fprintf(stderr, "\n");
        }
    }
    *out = reinterpret_cast<ckzg_b200_ctx*>(c);
    return RET_OK;
}

void ckzg_b200_ctx_destroy(ckzg_b200_ctx* ctx) { ctx_free(reinterpret_cast<Ctx*>(ctx)); }
int ckzg_b200_ctx_device(const ckzg_b200_ctx* ctx) { return reinterpret_cast<const Ctx*>(ctx)->device; }
// what the two lazily built fixed-base tables occupy right now (0 = not built yet / bucket form) and what the next build would choose
int ckzg_b200_ctx_table_info(const ckzg_b200_ctx* ctx, uint64_t out[6]) {
    if (!ctx || !out) return RET_BADARGS;
    const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
    out[0] = c->commit_table ? (uint64_t)fk_geom(c->commit_c).points_for(N_BLOB) * sizeof(G1Affine) : 0;
    out[1] = (uint64_t)(c->commit_table ? c->commit_c : 0);
    out[2] = c->fk_ready.load() ? (uint64_t)fk_geom(c->fk_c).table_points() * sizeof(G1Affine) : 0;
    out[3] = (uint64_t)(c->fk_ready.load() ? c->fk_c : 0);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    out[4] = (uint64_t)plan_commit_window();
    out[5] = (uint64_t)plan_fk_window();
    if (prev >= 0) cudaSetDevice(prev);
    return RET_OK;
}
int ckzg_b200_ctx_device_count(const ckzg_b200_ctx* ctx) {
    const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
    return c->peers.empty() ? 1 : (int)c->peers.size();
}
void ckzg_b200_set_caller_stream(void* cuda_stream) { caller_stream_tls() = (cudaStream_t)cuda_stream; }

uint64_t ckzg_b200_launch_count(const ckzg_b200_ctx* ctx) {
    const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
    uint64_t total = c->launches.load();
    for (size_t d = 1; d < c->peers.size(); d++) total += c->peers[d]->launches.load();
    return total;
}

}  // extern "C" (reopened below)
namespace kzg {
// n blobs -> n commitments.  Chunked so the sort lists (384 KiB/blob) stay bounded.
int commit_scalars_batch(Call& call, uint8_t* out_dev48, const uint8_t* d_scalars, bool big_endian, uint64_t n, int* d_bad) {
    Ctx* c = call.ctx;
    Launch L = call.launch();
    G1* d_res;
    TRY(call.alloc(&d_res, n));
    TRY(msm_direct_ensure(c));
    // off by default: for the 61 GB commitment table the two passes of level 1 are bound by random HBM reads and the
    // form only reaches parity with the XYZZ kernel (r02p: 30.2 ms against 30.4 ms per 1024 blobs); FK20 uses it
    static const int bam_min = getenv("CKZG_B200_AFFINE_MIN") ? atoi(getenv("CKZG_B200_AFFINE_MIN")) : 0;  // 0 = never
    if (c->commit_table && bam_min > 0 && n >= (uint64_t)bam_min) {
        // batches: pairwise affine additions with batched inversions (6 products per addition instead of 10)
        const uint64_t CH = 512;
        const uint64_t chunk = n < CH ? n : CH;
        uint8_t* ws;
        TRY(call.alloc(&ws, msm_affine_workspace_bytes(chunk, c->commit_c)));
        for (uint64_t off = 0; off < n; off += chunk) {
            const uint64_t m = (n - off < chunk) ? n - off : chunk;
            TRY(launch_msm_affine(L, d_res + off, d_scalars + off * BLOB_BYTES, big_endian, m, d_bad ? d_bad + off : nullptr, ws));
        }
        TRY(launch_g1_compress(L, out_dev48, d_res, n));
        return RET_OK;
    }
    if (c->commit_table) {  // direct table: no sort lists, no buckets
        uint8_t* ws;
        TRY(call.alloc(&ws, msm_direct_workspace_bytes(n)));
        TRY(launch_msm_direct(L, d_res, d_scalars, big_endian, n, d_bad, ws));
        TRY(launch_g1_compress(L, out_dev48, d_res, n));
        return RET_OK;
    }
    const uint64_t CHUNK = 2048;
    uint64_t chunk = n < CHUNK ? n : CHUNK;
    int parts = msm_pick_parts(chunk);
    uint8_t* ws;
    TRY(call.alloc(&ws, msm_workspace_bytes(chunk, parts)));
    for (uint64_t off = 0; off < n; off += chunk) {
        uint64_t m = (n - off < chunk) ? n - off : chunk;
        TRY(launch_msm(L, d_res + off, d_scalars + off * BLOB_BYTES, big_endian, m, c->msm_table, d_bad ? d_bad + off : nullptr, ws, parts));
    }
    TRY(launch_g1_compress(L, out_dev48, d_res, n));
    return RET_OK;
}
}  // namespace kzg
extern "C" {

int ckzg_b200_blob_to_kzg_commitment_batch(ckzg_b200_ctx* ctx, uint8_t* out, const uint8_t* blobs, uint64_t n, int mem, int* status) {
    if (!ctx || !out || !blobs) return RET_BADARGS;
    if (n == 0) return RET_OK;
    if (mem == CKZG_B200_HOST) {  // a context spanning devices: contiguous ranges of blobs, one per device
        Ctx* mc = reinterpret_cast<Ctx*>(ctx);
        const int parts = multi_parts(mc, n, 16);
        if (parts > 1)
            return multi_map(mc, n, parts, [&](ckzg_b200_ctx* dc, uint64_t f, uint64_t m) {
                return ckzg_b200_blob_to_kzg_commitment_batch(dc, out + 48 * f, blobs + f * BLOB_BYTES, m, mem, status ? status + f : nullptr);
            });
    }
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    const uint8_t* d_blobs;
    TRY(call.stage_in(&d_blobs, blobs, n * BLOB_BYTES, mem));
    int* d_bad;
    TRY(call.alloc(&d_bad, n));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, n * sizeof(int), call.stream));
    uint8_t* d_out;
    if (mem == CKZG_B200_DEVICE)
        d_out = out;
    else
        TRY(call.alloc(&d_out, n * 48));
    TRY(commit_scalars_batch(call, d_out, d_blobs, true, n, d_bad));
    std::vector<int> bad(n);
    KZG_CUDA_TRY(cudaMemcpyAsync(bad.data(), d_bad, n * sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    if (mem != CKZG_B200_DEVICE) KZG_CUDA_TRY(cudaMemcpyAsync(out, d_out, n * 48, cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    int rc = RET_OK;
    for (uint64_t i = 0; i < n; i++) {
        int s = bad[i] ? RET_BADARGS : RET_OK;
        if (status) status[i] = s;
        if (s && !rc) rc = s;
    }
    return rc;
}

void ckzg_b200_profile_enable(ckzg_b200_ctx* ctx, int on) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    std::lock_guard<std::mutex> g(c->prof.mu);
    c->prof.level = on < 0 ? 0 : (on > 2 ? 2 : on);
    c->prof.nk = 0;
    c->prof.call_ms = 0;
    c->prof.calls = 0;
}
int ckzg_b200_profile_dump(ckzg_b200_ctx* ctx, char* buf, size_t cap) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    std::lock_guard<std::mutex> g(c->prof.mu);
    int n = snprintf(buf, cap, "{\"calls\": %llu, \"call_ms\": %.6f, \"kernels\": {", (unsigned long long)c->prof.calls, c->prof.call_ms);
    for (int i = 0; i < c->prof.nk && n > 0 && (size_t)n < cap; i++)
        n += snprintf(buf + n, cap - n, "%s\"%s\": [%.6f, %llu]", i ? ", " : "", c->prof.names[i], c->prof.ms[i], (unsigned long long)c->prof.cnt[i]);
    if (n > 0 && (size_t)n < cap) n += snprintf(buf + n, cap - n, "}}");
    return n;
}

// measurement hook (tools/upload_probe.py): `reps` uploads of `bytes` from `host` through Call::stage_in -- the staged
// path for pageable memory, one DMA for pinned memory; mode 1 = cudaHostRegister + direct DMA + unregister instead
int ckzg_b200_debug_upload(ckzg_b200_ctx* ctx, const uint8_t* host, uint64_t bytes, int reps, int mode, double* ms_best) {
    if (!ctx || !host || !ms_best || reps < 1) return RET_BADARGS;
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        Call call(reinterpret_cast<Ctx*>(ctx));
        if (!call.ok) return RET_ERROR;
        uint8_t* d = nullptr;
        TRY(call.alloc(&d, bytes));
        KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
        const auto t0 = std::chrono::steady_clock::now();
        if (mode == 1) {
            KZG_CUDA_TRY(cudaHostRegister((void*)host, bytes, cudaHostRegisterDefault));
            KZG_CUDA_TRY(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, call.stream));
            KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
            KZG_CUDA_TRY(cudaHostUnregister((void*)host));
        } else {
            TRY(call.upload(d, host, bytes, call.stream));
            KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < best) best = ms;
    }
    *ms_best = best;
    return RET_OK;
}

int ckzg_b200_debug_placement(uint32_t* dev_buf) { return debug_set_placement_buffer(dev_buf); }
int ckzg_b200_debug_timers(uint32_t* dev_buf) { return debug_set_timer_buffer(dev_buf); }
int ckzg_b200_debug_pairing_probe(ckzg_b200_ctx* ctx, long long* ticks64, int* ok, const uint8_t* two_g1_48, int reps) {
    if (!ctx || !ticks64 || !ok || !two_g1_48 || reps < 1) return RET_BADARGS;
    return debug_pairing_probe(reinterpret_cast<Ctx*>(ctx), ticks64, ok, two_g1_48, reps);
}
int ckzg_b200_selftest_mulbench(int ilp, int iters, int blocks, int threads, float* ms_out) { return selftest_mulbench(ilp, iters, blocks, threads, ms_out); }
int ckzg_b200_selftest_field(int op, uint32_t* out, const uint32_t* a, const uint32_t* b, uint64_t n) { return selftest_field(op, out, a, b, n); }
int ckzg_b200_selftest_g1(int op, uint8_t* out48, int* ok_out, const uint8_t* p48, const uint32_t* k, const uint8_t* q48, uint64_t n) {
    return selftest_g1(op, out48, ok_out, p48, k, q48, n);
}

}  // extern "C"

// ---- entry points landing later this round (link-complete; fail loudly, never fall back) --------
extern "C" {
#ifndef KZG_HAVE_CELLS
#endif
}
