// Call-coalescing front end for the per-blob API: concurrent callers of
// blob_to_kzg_commitment / compute_*_proof / compute_cells_and_kzg_proofs / recover_cells_and_kzg_proofs
// (the reference is re-entrant and its consumers call it from many threads: bindings/go/main_test.go:957-970,
// bindings/rust/src/bindings/mod.rs:912) are merged into ONE batched engine call.
//
// Group commit: a caller that finds no batch running becomes the leader and runs at once -- a lone caller
// pays nothing; callers that arrive while a batch is running queue up and are taken together, in arrival
// order, by the next leader.  Up to `max_inflight` batches run concurrently so that one batch's upload
// overlaps another's kernels.
//
// Adaptive gather window (round 2): with many callers pure group commit alternates tiny and large batches (64 threads
// on compute_cells_and_kzg_proofs: 440 requests in 49 batches, average 9 -- 22 % of the batched throughput).  A leader
// that sees fewer waiters than the recent concurrency suggests (an average of "requests in flight or queued" per
// batch, halved per in-flight slot) now waits for them -- at most 1/16 of the last batch's duration and 500 us -- before
// it takes the batch.  A lone caller's estimate is 1, its target 1, its wait zero.
//
// Pure host C++ (no CUDA): tests/hostcheck compiles it with g++ against a mock executor.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

namespace kzg {

struct CoReq {
    const void* in[3] = {nullptr, nullptr, nullptr};
    void* out[2] = {nullptr, nullptr};
    uint64_t aux = 0;  // compatibility class: only requests with equal aux share a batch
    int rc = 0;
    bool done = false;
};

struct CombinerStats {
    uint64_t requests = 0, batches = 0, largest = 0;
};

class Combiner {
public:
    static constexpr int kUnserved = 3;  // C_KZG_MALLOC: what a caller sees if the executor could not run its batch
    explicit Combiner(size_t max_batch, int max_inflight = 2) : max_batch_(max_batch ? max_batch : 1), max_inflight_(max_inflight < 1 ? 1 : max_inflight) {}

    // Runs `r` through `run(std::vector<CoReq*>&)`, which must set rc of every request it is given.
    // Returns r.rc.  Blocks until r has been served (by this thread as a leader or by another one).
    template <class Run>
    int submit(CoReq& r, Run&& run) {
        std::unique_lock<std::mutex> lk(mu_);
        q_.push_back(&r);
        stats_.requests++;
        if (gathering_) cv_.notify_all();  // a leader is holding its batch open for arrivals like this one
        for (;;) {
            while (!r.done && (inflight_ >= max_inflight_ || q_.empty())) cv_.wait(lk);
            if (r.done) return r.rc;
            // gather window: hold the batch open while fewer callers wait than usually arrive together
            {
                const size_t want = (size_t)(conc_ / (double)max_inflight_);
                const size_t target = want < max_batch_ ? want : max_batch_;
                if (q_.size() < target && last_ns_ > 0) {
                    const uint64_t cap = last_ns_ / 16 < 500000 ? last_ns_ / 16 : 500000;
                    const auto deadline = std::chrono::steady_clock::now() + std::chrono::nanoseconds(cap);
                    gathering_++;
                    while (!r.done && q_.size() < target && inflight_ < max_inflight_)
                        if (cv_.wait_until(lk, deadline) == std::cv_status::timeout) break;
                    gathering_--;
                    if (r.done) return r.rc;
                    if (inflight_ >= max_inflight_ || q_.empty()) continue;  // another leader took it meanwhile
                }
            }
            // lead one batch: the longest prefix of the queue that shares the head's class
            std::vector<CoReq*> batch;
            const uint64_t cls = q_.front()->aux;
            while (!q_.empty() && batch.size() < max_batch_ && q_.front()->aux == cls) {
                batch.push_back(q_.front());
                q_.pop_front();
            }
            inflight_++;
            inflight_reqs_ += batch.size();
            stats_.batches++;
            if (batch.size() > stats_.largest) stats_.largest = batch.size();
            // callers alive right now: in flight (this batch included) + still queued
            conc_ = 0.75 * conc_ + 0.25 * (double)(inflight_reqs_ + q_.size());
            const auto t_run = std::chrono::steady_clock::now();
            lk.unlock();
            // a request is never reported as served unless the executor said so: an executor that throws
            // (allocation failure while gathering a batch) fails the batch instead of wedging the queue
            for (CoReq* b : batch) b->rc = kUnserved;
            try {
                run(batch);
            } catch (...) {
                for (CoReq* b : batch) b->rc = kUnserved;
            }
            lk.lock();
            inflight_--;
            inflight_reqs_ -= batch.size();
            last_ns_ = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_run).count();
            for (CoReq* b : batch) b->done = true;
            cv_.notify_all();
            if (r.done) return r.rc;
            // r was not in that batch (it sat behind max_batch others or another class): lead again / wait
        }
    }
    CombinerStats stats() {
        std::lock_guard<std::mutex> g(mu_);
        return stats_;
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<CoReq*> q_;
    size_t max_batch_;
    int max_inflight_;
    int inflight_ = 0;
    int gathering_ = 0;          // leaders currently holding a batch open
    size_t inflight_reqs_ = 0;   // requests inside running batches
    double conc_ = 1.0;          // running average of concurrently alive requests
    uint64_t last_ns_ = 0;       // duration of the batch that finished last
    CombinerStats stats_;
};

}  // namespace kzg
