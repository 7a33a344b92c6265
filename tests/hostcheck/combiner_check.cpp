// TEST-ONLY harness for the call-coalescing front end (c-kzg-4844_b200/csrc/combiner.h) with a mock
// executor: `threads` callers submit `per_thread` requests each; the executor sleeps `exec_us` per batch
// and answers request x with 3x + 1, or with rc = 1 when x is odd and `fail_odd` is set (fail_odd = 2: the executor
// throws on every batch that contains a multiple of 7 -- those callers must all see Combiner::kUnserved).
#include <atomic>
#include <chrono>
#include <thread>

#include "combiner.h"

using namespace kzg;

extern "C" int combiner_selftest(int threads, int per_thread, int max_batch, int max_inflight, int exec_us, int fail_odd, int classes, uint64_t* out4) {
    Combiner comb((size_t)max_batch, max_inflight);
    std::atomic<int> errors{0}, mixed{0}, oversize{0}, concurrent{0}, peak{0};
    auto run = [&](std::vector<CoReq*>& b) {
        int now = ++concurrent;
        int p = peak.load();
        while (now > p && !peak.compare_exchange_weak(p, now)) {}
        if ((int)b.size() > max_batch) oversize++;
        for (CoReq* r : b)
            if (r->aux != b[0]->aux) mixed++;
        std::this_thread::sleep_for(std::chrono::microseconds(exec_us));
        if (fail_odd == 2) {
            for (CoReq* r : b)
                if (*(const uint64_t*)r->in[0] % 7 == 0) {
                    --concurrent;
                    throw 1;
                }
        }
        for (CoReq* r : b) {
            uint64_t x = *(const uint64_t*)r->in[0];
            if (fail_odd == 1 && (x & 1)) {
                r->rc = 1;
            } else {
                *(uint64_t*)r->out[0] = 3 * x + 1;
                r->rc = 0;
            }
        }
        --concurrent;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] {
            for (int k = 0; k < per_thread; k++) {
                uint64_t x = (uint64_t)t * 1000003u + (uint64_t)k, y = 0;
                CoReq r;
                r.in[0] = &x;
                r.out[0] = &y;
                r.aux = classes > 1 ? (x % (uint64_t)classes) : 0;
                int rc = comb.submit(r, run);
                if (fail_odd == 2) {  // either served correctly or reported as unserved, never a silent wrong answer
                    if (!((rc == 0 && y == 3 * x + 1) || (rc == Combiner::kUnserved && y == 0))) errors++;
                    if (x % 7 == 0 && rc != Combiner::kUnserved) errors++;
                } else if (fail_odd == 1 && (x & 1)) {
                    if (rc != 1) errors++;
                } else if (rc != 0 || y != 3 * x + 1) {
                    errors++;
                }
            }
        });
    for (auto& t : th) t.join();
    CombinerStats s = comb.stats();
    out4[0] = s.requests;
    out4[1] = s.batches;
    out4[2] = s.largest;
    out4[3] = (uint64_t)peak.load();
    return errors.load() + 1000 * mixed.load() + 1000000 * oversize.load();
}
