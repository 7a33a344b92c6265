// Per-call plumbing shared by the api*.cu files.
#pragma once
#include <chrono>
#include <functional>
#include <vector>

#include "../../include/ckzg_b200.h"
#include "engine.h"

#include <stdlib.h>

#define TRY(expr)            \
    do {                     \
        int _rc = (expr);    \
        if (_rc) return _rc; \
    } while (0)

namespace kzg {

// Stream the calling thread produces its DEVICE-mode inputs on (ckzg_b200_set_caller_stream); nullptr = the
// legacy default stream, which the call streams below synchronise with implicitly.
cudaStream_t& caller_stream_tls();

// Per-call resources: a private stream and stream-ordered allocations (re-entrant: callers share a
// const context across threads, as the reference allows -- bindings/rust/src/bindings/mod.rs:912).
//
// Ordering against the caller (DEVICE-mode buffers): the call stream is a BLOCKING stream, so everything it
// does is ordered after work the caller enqueued earlier on the legacy default stream (torch's default);
// a caller that produces its buffers on another stream names it with ckzg_b200_set_caller_stream and the
// call stream waits for an event recorded there.  Side streams are forked from the call stream (fork()),
// so they inherit that ordering, and are joined back before the call's scratch is released on ANY exit path.
struct Call {
    Ctx* ctx;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    std::vector<cudaStream_t> sides;  // forked side streams, joined and destroyed by the destructor
    std::vector<std::pair<void*, size_t>> pinned;  // returned to the context's pool after the final sync
    int prev_device = -1;
    bool ok = false;
    bool pooled = true;           // streams come from / go back to the context's pool (not under profiling)
    bool profiling = false;       // whole-call begin/end events
    bool trace_kernels = false;   // per-kernel events (level 2)
    ProfTrace trace;
    ProfTrace timeline;           // level 1: completion times of side-stream work, relative to "begin"
    uint8_t* ring = nullptr;      // pinned staging ring of upload() (from the context's pool)
    cudaEvent_t ring_ev[4];
    std::chrono::steady_clock::time_point t_entry = std::chrono::steady_clock::now();
    // level-1 profiling only: HOST time since the call object was created (where the wall time outside the device
    // events goes: stream creation, enqueueing, the waits, teardown), reported as "host:..." entries of the dump
    void host_mark(const char* name) {
        if (!profiling || trace_kernels) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count();
        std::lock_guard<std::mutex> g(ctx->prof.mu);
        ctx->prof.add(name, ms);
    }

    explicit Call(Ctx* c) : ctx(c) {
        if (cudaGetDevice(&prev_device) != cudaSuccess) return;
        if (cudaSetDevice(c->device) != cudaSuccess) return;
        // level-1 profiling records timing events on every side stream; with pooled streams that run fell into the
        // late-event mode (18.6 ms per host batch against 13.5 unprofiled, profiles/e2e_pool_R3e.log), with streams of
        // its own it does not: profiled calls take fresh streams
        pooled = c->prof.level == 0;
        if (!(stream = c->stream_acquire(0, pooled))) return;
        if (cudaStream_t cs = caller_stream_tls()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return;
            cudaEventRecord(e, cs);
            cudaStreamWaitEvent(stream, e, 0);
            cudaEventDestroy(e);
        }
        ok = true;
        profiling = c->prof.level >= 1;
        trace_kernels = c->prof.level >= 2;
        if (profiling) mark("begin");
        host_mark("host:t_call_created");
    }
    ~Call() {
        if (stream) {
            join_sides();
            for (void* p : allocs) cudaFreeAsync(p, stream);
            if (profiling) mark("end");
            host_mark("host:t_teardown_enqueued");
            cudaStreamSynchronize(stream);
            host_mark("host:t_final_sync");
            if (profiling && trace.ev.size() >= 2) {
                std::lock_guard<std::mutex> g(ctx->prof.mu);
                float ms = 0;
                for (size_t i = 1; i < trace.ev.size(); i++) {
                    if (cudaEventElapsedTime(&ms, trace.ev[i - 1].first, trace.ev[i].first) == cudaSuccess) ctx->prof.add(trace.ev[i].second, ms);
                }
                if (cudaEventElapsedTime(&ms, trace.ev.front().first, trace.ev.back().first) == cudaSuccess) {
                    ctx->prof.call_ms += ms;
                    ctx->prof.calls++;
                }
                for (auto& t : timeline.ev)
                    if (cudaEventSynchronize(t.first) == cudaSuccess && cudaEventElapsedTime(&ms, trace.ev.front().first, t.first) == cudaSuccess) ctx->prof.add(t.second, ms);
            }
            if (ring)
                for (int i = 0; i < 4; i++)
                    if (ring_ev[i]) cudaEventDestroy(ring_ev[i]);
            for (auto& b : pinned) ctx->pin_release(b.first, b.second);
            for (auto& e : trace.ev) cudaEventDestroy(e.first);
            for (auto& e : timeline.ev) cudaEventDestroy(e.first);
            // last forked, first returned: the pool hands streams out from its end, so the next call's k-th fork gets
            // the stream this call's k-th fork had (the same hardware queue for the same role, call after call)
            for (size_t i = sides.size(); i-- > 0;) ctx->stream_release(1, sides[i], pooled);
            ctx->stream_release(0, stream, pooled);
            host_mark("host:t_exit");
        }
        if (prev_device >= 0) cudaSetDevice(prev_device);
    }
    // A non-blocking side stream that starts after everything enqueued so far on the call stream.
    // nullptr on failure.  Owned by the call: joined (join_sides) and destroyed with it.
    cudaStream_t fork() {
        cudaStream_t sd = nullptr;
        cudaEvent_t e = nullptr;
        if (!(sd = ctx->stream_acquire(1, pooled))) return nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            ctx->stream_release(1, sd, pooled);
            return nullptr;
        }
        cudaEventRecord(e, stream);
        cudaStreamWaitEvent(sd, e, 0);
        cudaEventDestroy(e);
        sides.push_back(sd);
        return sd;
    }
    // the call stream waits for everything enqueued on the side streams (idempotent; cheap when they are idle)
    void join_sides() {
        for (cudaStream_t sd : sides) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                cudaStreamSynchronize(sd);
                continue;
            }
            cudaEventRecord(e, sd);
            cudaStreamWaitEvent(stream, e, 0);
            cudaEventDestroy(e);
        }
    }
    template <class T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e != cudaSuccess) {
            note_cuda_error(e, __FILE__, __LINE__);
            return e == cudaErrorMemoryAllocation ? RET_MALLOC : RET_ERROR;
        }
        allocs.push_back(p);
        *out = (T*)p;
        return RET_OK;
    }
    // pinned host scratch, valid until the call object dies
    int pin(uint8_t** out, size_t bytes) {
        size_t cap = 0;
        void* p = ctx->pin_acquire(bytes, &cap);
        if (!p) return RET_MALLOC;
        pinned.push_back({p, cap});
        *out = (uint8_t*)p;
        return RET_OK;
    }
    // input in `mem` space -> device pointer (copy if host)
    int stage_in(const uint8_t** dev, const uint8_t* src, size_t bytes, int mem) {
        if (mem == CKZG_B200_DEVICE) {
            *dev = src;
            return RET_OK;
        }
        uint8_t* d;
        int rc = alloc(&d, bytes);
        if (rc) return rc;
        rc = upload(d, src, bytes, stream);
        if (rc) return rc;
        *dev = d;
        return RET_OK;
    }
    // Host -> device copy of `bytes` on stream `st` (ordered like a cudaMemcpyAsync there).  Pinned / registered
    // sources go straight to the copy engine.  PAGEABLE sources of a few MB or more (what the frozen API's callers
    // pass: Go slices, Python bytes) are copied by the context's host threads into a ring of pinned slots, each
    // sent while the next fills (hostpool.h); the function returns once the last slot is queued, the data has left
    // the caller's buffer by then.
    int upload(void* dst, const void* src, size_t bytes, cudaStream_t st) {
        static const size_t kMin = getenv("CKZG_B200_STAGE_MIN") ? (size_t)atoll(getenv("CKZG_B200_STAGE_MIN")) : (size_t)(4u << 20);
        if (bytes < kMin || !host_ptr_is_pageable(src)) {
            KZG_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
            return RET_OK;
        }
        // CKZG_B200_STAGE_SLOT_MB: size of one of the four pinned staging slots (default 8 MiB)
        static const size_t SLOT = ((getenv("CKZG_B200_STAGE_SLOT_MB") && atoi(getenv("CKZG_B200_STAGE_SLOT_MB")) > 0) ? (size_t)atoi(getenv("CKZG_B200_STAGE_SLOT_MB")) : 8u) << 20;
        constexpr int NSLOT = 4;
        if (!ring) {
            TRY(pin(&ring, SLOT * NSLOT));
            for (int i = 0; i < NSLOT; i++) ring_ev[i] = nullptr;
        }
        HostPool* pool = ctx->host_pool();
        const int parts = pool->threads() + 1;
        size_t off = 0;
        for (int k = 0; off < bytes; k++) {
            const int slot = k % NSLOT;
            const size_t len = bytes - off < SLOT ? bytes - off : SLOT;
            if (ring_ev[slot]) KZG_CUDA_TRY(cudaEventSynchronize(ring_ev[slot]));  // the copy engine is done with this slot
            else KZG_CUDA_TRY(cudaEventCreateWithFlags(&ring_ev[slot], cudaEventDisableTiming));
            uint8_t* s8 = ring + (size_t)slot * SLOT;
            const uint8_t* from = (const uint8_t*)src + off;
            const size_t per = ((len + parts - 1) / parts + 4095) & ~(size_t)4095;
            pool->run(parts, [&](int i) {
                const size_t a = (size_t)i * per;
                if (a < len) memcpy(s8 + a, from + a, len - a < per ? len - a : per);
            });
            KZG_CUDA_TRY(cudaMemcpyAsync((uint8_t*)dst + off, s8, len, cudaMemcpyHostToDevice, st));
            KZG_CUDA_TRY(cudaEventRecord(ring_ev[slot], st));
            off += len;
        }
        return RET_OK;
    }
    Launch launch() { return Launch{ctx, stream, trace_kernels ? &trace : nullptr}; }
    Launch launch_on(cudaStream_t s) { return Launch{ctx, s, nullptr}; }
    // level-1 profiling only: when did the work enqueued so far on stream `s` finish, relative to "begin"
    void mark_on(cudaStream_t s, const char* name) {
        if (!profiling || trace_kernels) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) == cudaSuccess) {
            cudaEventRecord(e, s);
            timeline.ev.push_back({e, name});
        }
    }
    void mark(const char* name) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) == cudaSuccess) {
            cudaEventRecord(e, stream);
            trace.ev.push_back({e, name});
        }
    }
};



// in-library multi-device fan-out (multi.cu)
int multi_device_count(const Ctx* c);
int multi_parts(const Ctx* c, uint64_t n, uint64_t min_per_part);
bool multi_inside_fanout();
int multi_map(Ctx* c, uint64_t n, int parts, const std::function<int(ckzg_b200_ctx*, uint64_t, uint64_t)>& fn);

// call-coalescing front end of the per-blob entry points (coalesce.cu)
void coalescer_create(Ctx* c);
void coalescer_destroy(Ctx* c);

// n x 4096 scalars (wire blobs or plain limbs) -> n compressed commitments (api.cu)
int commit_scalars_batch(Call& call, uint8_t* out_dev48, const uint8_t* d_scalars, bool big_endian, uint64_t n, int* d_bad);

}  // namespace kzg
