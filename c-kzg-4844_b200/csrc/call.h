// Per-call plumbing shared by the api*.cu files.
#pragma once
#include <vector>

#include "../../include/ckzg_b200.h"
#include "engine.h"

namespace kzg {

// Per-call resources: a private stream and stream-ordered allocations (re-entrant: callers share a
// const context across threads, as the reference allows -- bindings/rust/src/bindings/mod.rs:912).
struct Call {
    Ctx* ctx;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    std::vector<std::pair<void*, size_t>> pinned;  // returned to the context's pool after the final sync
    int prev_device = -1;
    bool ok = false;
    bool profiling = false;       // whole-call begin/end events
    bool trace_kernels = false;   // per-kernel events (level 2)
    ProfTrace trace;
    ProfTrace timeline;           // level 1: completion times of side-stream work, relative to "begin"

    explicit Call(Ctx* c) : ctx(c) {
        if (cudaGetDevice(&prev_device) != cudaSuccess) return;
        if (cudaSetDevice(c->device) != cudaSuccess) return;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return;
        ok = true;
        profiling = c->prof.level >= 1;
        trace_kernels = c->prof.level >= 2;
        if (profiling) mark("begin");
    }
    ~Call() {
        if (stream) {
            for (void* p : allocs) cudaFreeAsync(p, stream);
            if (profiling) mark("end");
            cudaStreamSynchronize(stream);
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
            for (auto& b : pinned) ctx->pin_release(b.first, b.second);
            for (auto& e : trace.ev) cudaEventDestroy(e.first);
            for (auto& e : timeline.ev) cudaEventDestroy(e.first);
            cudaStreamDestroy(stream);
        }
        if (prev_device >= 0) cudaSetDevice(prev_device);
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
        KZG_CUDA_TRY(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, stream));
        *dev = d;
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

#define TRY(expr)            \
    do {                     \
        int _rc = (expr);    \
        if (_rc) return _rc; \
    } while (0)


// call-coalescing front end of the per-blob entry points (coalesce.cu)
void coalescer_create(Ctx* c);
void coalescer_destroy(Ctx* c);

// n x 4096 scalars (wire blobs or plain limbs) -> n compressed commitments (api.cu)
int commit_scalars_batch(Call& call, uint8_t* out_dev48, const uint8_t* d_scalars, bool big_endian, uint64_t n, int* d_bad);

}  // namespace kzg
