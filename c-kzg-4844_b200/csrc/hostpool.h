// Host-side helpers of the C-ABI layer: a small persistent worker pool (parallel memcpy into pinned staging,
// per-device / per-sub-batch fan-out) and the staged upload of PAGEABLE caller buffers.
//
// Why: the frozen API takes plain host pointers -- Go slices (bindings/go/main.go:441-461), Python bytes
// (bindings/python/ckzg_wrap.c) -- i.e. pageable memory.  cudaMemcpyAsync from pageable memory is staged by the
// driver through one thread and one small bounce buffer (~10 GB/s, and it blocks the caller); a 537 MB batch of
// blobs then costs several times the 9.8 ms its PCIe transfer needs.  Here the caller's bytes are copied by several
// host threads into a ring of pinned slots and each slot is sent by the copy engine while the next one fills.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace kzg {

// parallel_for over a fixed set of persistent threads.  run(n, fn): fn(i) for i in [0, n), the caller takes part;
// returns when all are done.  Re-entrant across callers (jobs queue up behind one lock; a caller whose job is queued
// works on it itself, so nested or concurrent use cannot deadlock).
class HostPool {
  public:
    explicit HostPool(int threads) {
        for (int i = 0; i < threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int threads() const { return (int)workers_.size(); }
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        Job job{&fn, n, 0, 0};
        {
            std::lock_guard<std::mutex> g(mu_);
            jobs_.push_back(&job);
        }
        cv_.notify_all();
        work_on(job);  // the caller helps; when it runs out of indices it waits for the stragglers
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return job.done == job.n; });
    }

  private:
    struct Job {
        const std::function<void(int)>* fn;
        int n, next, done;
    };
    void work_on(Job& job) {
        for (;;) {
            int i;
            {
                std::lock_guard<std::mutex> g(mu_);
                if (job.next >= job.n) {
                    for (size_t k = 0; k < jobs_.size(); k++)
                        if (jobs_[k] == &job) {
                            jobs_.erase(jobs_.begin() + k);
                            break;
                        }
                    return;
                }
                i = job.next++;
            }
            (*job.fn)(i);
            {
                std::lock_guard<std::mutex> g(mu_);
                job.done++;
                if (job.done == job.n) done_cv_.notify_all();
            }
        }
    }
    void loop() {
        for (;;) {
            Job* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (stop_) return;
                job = jobs_.front();
                if (job->next >= job->n) {  // exhausted, its owner will remove it
                    jobs_.erase(jobs_.begin());
                    continue;
                }
            }
            work_on_shared(*job);
        }
    }
    // worker side: take indices while the job is still queued (the owner may have finished and left)
    void work_on_shared(Job& job) {
        for (;;) {
            int i;
            const std::function<void(int)>* fn;
            {
                std::lock_guard<std::mutex> g(mu_);
                bool queued = false;
                for (Job* j : jobs_) queued = queued || (j == &job);
                if (!queued || job.next >= job.n) return;
                i = job.next++;
                fn = job.fn;
            }
            (*fn)(i);
            {
                std::lock_guard<std::mutex> g(mu_);
                job.done++;
                if (job.done == job.n) done_cv_.notify_all();
            }
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<Job*> jobs_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
};

// CKZG_B200_HOST_THREADS (default: min(8, hardware threads / visible devices), at least 2).  Measured on a 16-core
// B200 host (tools/upload_probe.py, profiles/R2_summary.md): staging 512 MiB of pageable memory ALONE runs at 23.7 /
// 35.9 / 42.1 GB/s with 2 / 8 / 16 threads and 8 MiB slots, against 55.6 GB/s for pinned memory and 18.4 GB/s for
// cudaHostRegister + direct DMA + unregister.  Inside a real call 16 threads on 16 cores lose to 8 (e2e_pageable 27 ms
// against 19-22 ms at 4096 blobs: the caller's launch thread and the driver's own threads need cores too), so 8 it is.
inline int host_threads_default() {
    if (const char* env = getenv("CKZG_B200_HOST_THREADS")) {
        const int v = atoi(env);
        if (v >= 1 && v <= 64) return v;
    }
    int ndev = 1;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) ndev = 1;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 4;
    int t = hw / ndev;
    return t > 8 ? 8 : (t < 2 ? 2 : t);
}

// true if `p` is ordinary pageable host memory (not cudaHostAlloc / cudaHostRegister memory)
inline bool host_ptr_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace kzg
