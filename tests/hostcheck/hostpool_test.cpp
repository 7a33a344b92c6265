// CPU test of csrc/hostpool.h (the worker pool behind staged uploads and sub-batch fan-out): every index runs exactly
// once, nested and concurrent run() calls neither deadlock nor lose work.  Built and run by tests/test_hostpool.py.
#include "hostpool.h"
#include <atomic>
#include <cstdio>
using namespace kzg;
int main() {
    HostPool pool(5);
    std::atomic<long> sum{0};
    for (int rep = 0; rep < 2000; rep++) {
        int n = 1 + rep % 17;
        std::atomic<int> cnt{0};
        pool.run(n, [&](int i) { cnt++; sum += i; });
        if (cnt != n) { printf("FAIL %d %d\n", (int)cnt, n); return 1; }
    }
    // concurrent callers
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (int t = 0; t < 6; t++) th.emplace_back([&] {
        for (int rep = 0; rep < 500; rep++) {
            std::atomic<int> cnt{0};
            pool.run(9, [&](int i) { cnt++; if (i == 3) pool.run(3, [&](int) { cnt++; }); });
            if (cnt != 12) bad++;
        }
    });
    for (auto& t : th) t.join();
    printf("ok bad=%d sum=%ld\n", (int)bad, (long)sum);
    return bad != 0;
}
