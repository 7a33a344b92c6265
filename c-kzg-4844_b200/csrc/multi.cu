// In-library multi-device fan-out: a context created with CKZG_B200_DEVICES=0,1,...
// owns one ordinary context per device; the batched entry points of the C ABI cut a HOST-memory batch into
// contiguous ranges, one host thread per range drives its device through the SAME single-device entry point, and the
// per-blob status codes / outputs land in the caller's arrays at their own offsets.  Nothing is exchanged between
// devices for the per-blob paths (commitments, proofs, cells, recovery: replicas only); verify_blob_kzg_proof_batch
// has its own sharded form with one exchange through pinned host memory (api_verify.cu verify_blob_batch_multi);
// verify_cell_kzg_proof_batch is cut into independently verified sub-batches (ranges of cells, own challenge each,
// verdicts AND-ed: what the reference's parallel benchmark does, bindings/go/main_test.go:1037-1101).
// This is what makes a second GPU reachable from a binding that only knows the frozen API.
#include <thread>

#include "call.h"

namespace kzg {

static thread_local bool tl_inside_fanout = false;

int multi_device_count(const Ctx* c) { return c->peers.empty() ? 1 : (int)c->peers.size(); }

bool multi_inside_fanout() { return tl_inside_fanout; }

// parts = min(devices, n / min_per_part); 1 means: do not shard (single device, nested call, small batch)
int multi_parts(const Ctx* c, uint64_t n, uint64_t min_per_part) {
    if (tl_inside_fanout || c->peers.size() < 2 || min_per_part == 0) return 1;
    const uint64_t by_size = n / min_per_part;
    const uint64_t d = c->peers.size();
    const uint64_t p = by_size < d ? by_size : d;
    return p < 2 ? 1 : (int)p;
}

// fn(context of the part's device, first, count) for `parts` contiguous ranges of [0, n); part p runs on device
// p * devices / parts.  Returns the first non-OK code in range order (what a sequential pass would have reported).
int multi_map(Ctx* c, uint64_t n, int parts, const std::function<int(ckzg_b200_ctx*, uint64_t, uint64_t)>& fn) {
    const int D = multi_device_count(c);
    std::vector<int> rc(parts, RET_OK);
    auto body = [&](int p) {
        const uint64_t first = n * (uint64_t)p / (uint64_t)parts, end = n * (uint64_t)(p + 1) / (uint64_t)parts;
        Ctx* dc = c->peers.empty() ? c : c->peers[(size_t)p * D / parts];
        const bool was = tl_inside_fanout;
        tl_inside_fanout = true;
        rc[p] = end > first ? fn(reinterpret_cast<ckzg_b200_ctx*>(dc), first, end - first) : RET_OK;
        tl_inside_fanout = was;
    };
    std::vector<std::thread> th;
    for (int p = 1; p < parts; p++) th.emplace_back(body, p);
    body(0);
    for (auto& t : th) t.join();
    for (int p = 0; p < parts; p++)
        if (rc[p]) return rc[p];
    return RET_OK;
}

}  // namespace kzg
