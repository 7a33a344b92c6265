// Per-blob entry points of the C ABI that merge concurrent callers into batched engine calls
// (combiner.h).  The frozen API in src/ckzg.c calls these; a batch of one
// goes straight through with the caller's own pointers, larger batches are gathered into pinned
// staging buffers, run through the *_batch entry points with per-blob status, and scattered back.
// Host-side plumbing only: every operation still runs in the kernels behind the *_batch calls.
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "call.h"
#include "cells.h"
#include "combiner.h"

namespace kzg {

struct Coalescer {
    Combiner commit{256};
    Combiner blob_proof{256};
    Combiner kzg_proof{256};
    Combiner cells{64};
    Combiner recover{64};
    bool enabled = true;
};

void coalescer_create(Ctx* c) {
    Coalescer* co = new (std::nothrow) Coalescer();
    if (co) {
        const char* env = getenv("CKZG_B200_COALESCE");
        if (env && env[0] == '0') co->enabled = false;
    }
    c->coalescer = co;
}
void coalescer_destroy(Ctx* c) {
    delete reinterpret_cast<Coalescer*>(c->coalescer);
    c->coalescer = nullptr;
}

namespace {

// pinned staging block of the context's pool, released on scope exit
struct Pinned {
    Ctx* c;
    void* p = nullptr;
    size_t cap = 0;
    Pinned(Ctx* ctx, size_t bytes) : c(ctx) { p = bytes ? c->pin_acquire(bytes, &cap) : nullptr; }
    ~Pinned() {
        if (p) c->pin_release(p, cap);
    }
    uint8_t* u8() const { return (uint8_t*)p; }
};

void fail_all(std::vector<CoReq*>& b, int rc) {
    for (CoReq* r : b) r->rc = rc;
}
// per-blob status of a batched call: entries the engine never reached keep the call's return code
void spread_status(std::vector<CoReq*>& b, const std::vector<int>& status, int rc) {
    for (size_t i = 0; i < b.size(); i++) b[i]->rc = status[i] >= 0 ? status[i] : (rc ? rc : RET_ERROR);
}

Coalescer* coalescer_of(ckzg_b200_ctx* ctx) {
    Coalescer* co = reinterpret_cast<Coalescer*>(reinterpret_cast<Ctx*>(ctx)->coalescer);
    return (co && co->enabled) ? co : nullptr;
}

}  // namespace
}  // namespace kzg

using namespace kzg;

extern "C" {

int ckzg_b200_blob_to_kzg_commitment_coalesced(ckzg_b200_ctx* ctx, uint8_t* out48, const uint8_t* blob) {
    if (!ctx || !out48 || !blob) return RET_BADARGS;
    Coalescer* co = coalescer_of(ctx);
    if (!co) return ckzg_b200_blob_to_kzg_commitment_batch(ctx, out48, blob, 1, CKZG_B200_HOST, nullptr);
    CoReq r;
    r.in[0] = blob;
    r.out[0] = out48;
    return co->commit.submit(r, [ctx](std::vector<CoReq*>& b) {
        const size_t n = b.size();
        if (n == 1) {
            b[0]->rc = ckzg_b200_blob_to_kzg_commitment_batch(ctx, (uint8_t*)b[0]->out[0], (const uint8_t*)b[0]->in[0], 1, CKZG_B200_HOST, nullptr);
            return;
        }
        Pinned in(reinterpret_cast<Ctx*>(ctx), n * BLOB_BYTES), out(reinterpret_cast<Ctx*>(ctx), n * 48);
        if (!in.p || !out.p) return fail_all(b, RET_MALLOC);
        for (size_t i = 0; i < n; i++) memcpy(in.u8() + i * BLOB_BYTES, b[i]->in[0], BLOB_BYTES);
        std::vector<int> status(n, -1);
        int rc = ckzg_b200_blob_to_kzg_commitment_batch(ctx, out.u8(), in.u8(), n, CKZG_B200_HOST, status.data());
        for (size_t i = 0; i < n; i++)
            if (status[i] == RET_OK) memcpy(b[i]->out[0], out.u8() + 48 * i, 48);
        spread_status(b, status, rc);
    });
}

int ckzg_b200_compute_blob_kzg_proof_coalesced(ckzg_b200_ctx* ctx, uint8_t* proof48, const uint8_t* blob, const uint8_t* commitment48) {
    if (!ctx || !proof48 || !blob || !commitment48) return RET_BADARGS;
    Coalescer* co = coalescer_of(ctx);
    if (!co) return ckzg_b200_compute_blob_kzg_proof_batch(ctx, proof48, blob, commitment48, 1, CKZG_B200_HOST, nullptr);
    CoReq r;
    r.in[0] = blob;
    r.in[1] = commitment48;
    r.out[0] = proof48;
    return co->blob_proof.submit(r, [ctx](std::vector<CoReq*>& b) {
        const size_t n = b.size();
        if (n == 1) {
            b[0]->rc = ckzg_b200_compute_blob_kzg_proof_batch(ctx, (uint8_t*)b[0]->out[0], (const uint8_t*)b[0]->in[0], (const uint8_t*)b[0]->in[1], 1, CKZG_B200_HOST, nullptr);
            return;
        }
        Pinned in(reinterpret_cast<Ctx*>(ctx), n * (BLOB_BYTES + 48)), out(reinterpret_cast<Ctx*>(ctx), n * 48);
        if (!in.p || !out.p) return fail_all(b, RET_MALLOC);
        uint8_t* cms = in.u8() + n * BLOB_BYTES;
        for (size_t i = 0; i < n; i++) {
            memcpy(in.u8() + i * BLOB_BYTES, b[i]->in[0], BLOB_BYTES);
            memcpy(cms + 48 * i, b[i]->in[1], 48);
        }
        std::vector<int> status(n, -1);
        int rc = ckzg_b200_compute_blob_kzg_proof_batch(ctx, out.u8(), in.u8(), cms, n, CKZG_B200_HOST, status.data());
        for (size_t i = 0; i < n; i++)
            if (status[i] == RET_OK) memcpy(b[i]->out[0], out.u8() + 48 * i, 48);
        spread_status(b, status, rc);
    });
}

int ckzg_b200_compute_kzg_proof_coalesced(ckzg_b200_ctx* ctx, uint8_t* proof48, uint8_t* y32, const uint8_t* blob, const uint8_t* z32) {
    if (!ctx || !proof48 || !y32 || !blob || !z32) return RET_BADARGS;
    Coalescer* co = coalescer_of(ctx);
    if (!co) return ckzg_b200_compute_kzg_proof_batch(ctx, proof48, y32, blob, z32, 1, CKZG_B200_HOST, nullptr);
    CoReq r;
    r.in[0] = blob;
    r.in[1] = z32;
    r.out[0] = proof48;
    r.out[1] = y32;
    return co->kzg_proof.submit(r, [ctx](std::vector<CoReq*>& b) {
        const size_t n = b.size();
        if (n == 1) {
            b[0]->rc = ckzg_b200_compute_kzg_proof_batch(ctx, (uint8_t*)b[0]->out[0], (uint8_t*)b[0]->out[1], (const uint8_t*)b[0]->in[0], (const uint8_t*)b[0]->in[1], 1,
                                                         CKZG_B200_HOST, nullptr);
            return;
        }
        Pinned in(reinterpret_cast<Ctx*>(ctx), n * (BLOB_BYTES + 32)), out(reinterpret_cast<Ctx*>(ctx), n * 80);
        if (!in.p || !out.p) return fail_all(b, RET_MALLOC);
        uint8_t* zs = in.u8() + n * BLOB_BYTES;
        uint8_t* ys = out.u8() + n * 48;
        for (size_t i = 0; i < n; i++) {
            memcpy(in.u8() + i * BLOB_BYTES, b[i]->in[0], BLOB_BYTES);
            memcpy(zs + 32 * i, b[i]->in[1], 32);
        }
        std::vector<int> status(n, -1);
        int rc = ckzg_b200_compute_kzg_proof_batch(ctx, out.u8(), ys, in.u8(), zs, n, CKZG_B200_HOST, status.data());
        for (size_t i = 0; i < n; i++)
            if (status[i] == RET_OK) {
                memcpy(b[i]->out[0], out.u8() + 48 * i, 48);
                memcpy(b[i]->out[1], ys + 32 * i, 32);
            }
        spread_status(b, status, rc);
    });
}

int ckzg_b200_compute_cells_and_kzg_proofs_coalesced(ckzg_b200_ctx* ctx, uint8_t* cells, uint8_t* proofs, const uint8_t* blob) {
    if (!ctx || !blob) return RET_BADARGS;
    if (!cells && !proofs) return RET_BADARGS;  // eip7594.c:72-74
    Coalescer* co = coalescer_of(ctx);
    if (!co) return ckzg_b200_compute_cells_and_kzg_proofs_batch(ctx, cells, proofs, blob, 1, CKZG_B200_HOST, nullptr);
    CoReq r;
    r.in[0] = blob;
    r.out[0] = cells;
    r.out[1] = proofs;
    r.aux = proofs ? 1 : 0;  // proof requests (FK20, ~99 % of the cost) do not hold up cells-only callers
    return co->cells.submit(r, [ctx](std::vector<CoReq*>& b) {
        const size_t n = b.size();
        if (n == 1) {
            b[0]->rc = ckzg_b200_compute_cells_and_kzg_proofs_batch(ctx, (uint8_t*)b[0]->out[0], (uint8_t*)b[0]->out[1], (const uint8_t*)b[0]->in[0], 1, CKZG_B200_HOST, nullptr);
            return;
        }
        bool want_cells = false;
        const bool want_proofs = b[0]->aux != 0;
        for (CoReq* q : b) want_cells |= q->out[0] != nullptr;
        Ctx* c = reinterpret_cast<Ctx*>(ctx);
        Pinned in(c, n * BLOB_BYTES), oc(c, want_cells ? n * 2 * BLOB_BYTES : 0), op(c, want_proofs ? n * CELLS_EXT * 48 : 0);
        if (!in.p || (want_cells && !oc.p) || (want_proofs && !op.p)) return fail_all(b, RET_MALLOC);
        for (size_t i = 0; i < n; i++) memcpy(in.u8() + i * BLOB_BYTES, b[i]->in[0], BLOB_BYTES);
        std::vector<int> status(n, -1);
        int rc = ckzg_b200_compute_cells_and_kzg_proofs_batch(ctx, oc.u8(), op.u8(), in.u8(), n, CKZG_B200_HOST, status.data());
        for (size_t i = 0; i < n; i++)
            if (status[i] == RET_OK) {
                if (b[i]->out[0]) memcpy(b[i]->out[0], oc.u8() + i * 2 * BLOB_BYTES, 2 * BLOB_BYTES);
                if (b[i]->out[1]) memcpy(b[i]->out[1], op.u8() + i * CELLS_EXT * 48, CELLS_EXT * 48);
            }
        spread_status(b, status, rc);
    });
}

int ckzg_b200_recover_cells_and_kzg_proofs_coalesced(ckzg_b200_ctx* ctx, uint8_t* recovered_cells, uint8_t* recovered_proofs, const uint64_t* cell_indices, const uint8_t* cells,
                                                     uint64_t num_cells) {
    if (!ctx || !recovered_cells || !cell_indices || !cells) return RET_BADARGS;
    // eip7594.c:191-213: a request with bad counts / indices never enters a batch (it would fail its neighbours)
    if (num_cells > CELLS_EXT || num_cells < CELLS_EXT / 2) return RET_BADARGS;
    for (uint64_t i = 0; i < num_cells; i++) {
        if (cell_indices[i] >= CELLS_EXT) return RET_BADARGS;
        if (i > 0 && cell_indices[i] <= cell_indices[i - 1]) return RET_BADARGS;
    }
    Coalescer* co = coalescer_of(ctx);
    if (!co) return ckzg_b200_recover_cells_and_kzg_proofs_batch(ctx, recovered_cells, recovered_proofs, cell_indices, cells, num_cells, 1, CKZG_B200_HOST, nullptr);
    CoReq r;
    r.in[0] = cells;
    r.in[1] = cell_indices;
    r.out[0] = recovered_cells;
    r.out[1] = recovered_proofs;
    r.aux = num_cells * 2 + (recovered_proofs ? 1 : 0);  // the batched entry takes one cell count per batch
    return co->recover.submit(r, [ctx](std::vector<CoReq*>& b) {
        const size_t n = b.size();
        const uint64_t nc = b[0]->aux >> 1;
        const bool want_proofs = (b[0]->aux & 1) != 0;
        if (n == 1) {
            b[0]->rc = ckzg_b200_recover_cells_and_kzg_proofs_batch(ctx, (uint8_t*)b[0]->out[0], (uint8_t*)b[0]->out[1], (const uint64_t*)b[0]->in[1], (const uint8_t*)b[0]->in[0], nc, 1,
                                                                    CKZG_B200_HOST, nullptr);
            return;
        }
        Ctx* c = reinterpret_cast<Ctx*>(ctx);
        Pinned in(c, n * nc * CELL_BYTES), oc(c, n * 2 * BLOB_BYTES), op(c, want_proofs ? n * CELLS_EXT * 48 : 0);
        if (!in.p || !oc.p || (want_proofs && !op.p)) return fail_all(b, RET_MALLOC);
        std::vector<uint64_t> idx(n * nc);
        for (size_t i = 0; i < n; i++) {
            memcpy(in.u8() + i * nc * CELL_BYTES, b[i]->in[0], nc * CELL_BYTES);
            memcpy(idx.data() + i * nc, b[i]->in[1], nc * sizeof(uint64_t));
        }
        std::vector<int> status(n, -1);
        int rc = ckzg_b200_recover_cells_and_kzg_proofs_batch(ctx, oc.u8(), op.u8(), idx.data(), in.u8(), nc, n, CKZG_B200_HOST, status.data());
        for (size_t i = 0; i < n; i++)
            if (status[i] == RET_OK) {
                memcpy(b[i]->out[0], oc.u8() + i * 2 * BLOB_BYTES, 2 * BLOB_BYTES);
                if (b[i]->out[1]) memcpy(b[i]->out[1], op.u8() + i * CELLS_EXT * 48, CELLS_EXT * 48);
            }
        spread_status(b, status, rc);
    });
}

// {requests, batches, largest batch} of the five combiners, in the order commit, blob_proof, kzg_proof, cells, recover
int ckzg_b200_coalesce_stats(ckzg_b200_ctx* ctx, uint64_t out15[15]) {
    if (!ctx || !out15) return RET_BADARGS;
    Coalescer* co = reinterpret_cast<Coalescer*>(reinterpret_cast<Ctx*>(ctx)->coalescer);
    if (!co) return RET_ERROR;
    Combiner* all[5] = {&co->commit, &co->blob_proof, &co->kzg_proof, &co->cells, &co->recover};
    for (int i = 0; i < 5; i++) {
        CombinerStats s = all[i]->stats();
        out15[3 * i] = s.requests;
        out15[3 * i + 1] = s.batches;
        out15[3 * i + 2] = s.largest;
    }
    return RET_OK;
}

// 0 = every call runs alone (as before), 1 = merge concurrent callers (default; env CKZG_B200_COALESCE=0 disables)
int ckzg_b200_coalesce_enable(ckzg_b200_ctx* ctx, int on) {
    if (!ctx) return RET_BADARGS;
    Coalescer* co = reinterpret_cast<Coalescer*>(reinterpret_cast<Ctx*>(ctx)->coalescer);
    if (!co) return RET_ERROR;
    co->enabled = on != 0;
    return RET_OK;
}

// Measurement hook (bench.py): `threads` native host threads call the per-blob entry point `op`
// (0 = blob_to_kzg_commitment, 1 = compute_cells_and_kzg_proofs) `reps` times each on blobs taken from a host
// array, exactly as concurrent users of the frozen API do; *seconds = wall time of the whole run.
int ckzg_b200_bench_per_blob_callers(ckzg_b200_ctx* ctx, int op, int threads, int reps, const uint8_t* blobs, uint64_t n_blobs, double* seconds) {
    if (!ctx || !blobs || !seconds || threads < 1 || reps < 1 || n_blobs == 0 || op < 0 || op > 1) return RET_BADARGS;
    std::vector<int> rcs((size_t)threads, RET_OK);
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; t++)
        th.emplace_back([=, &rcs] {
            std::vector<uint8_t> cells(op == 1 ? 2 * (size_t)BLOB_BYTES : 0), proofs(op == 1 ? CELLS_EXT * 48 : 48);
            for (int k = 0; k < reps; k++) {
                const uint8_t* blob = blobs + (((uint64_t)t * (uint64_t)reps + (uint64_t)k) % n_blobs) * BLOB_BYTES;
                int rc = op == 0 ? ckzg_b200_blob_to_kzg_commitment_coalesced(ctx, proofs.data(), blob)
                                 : ckzg_b200_compute_cells_and_kzg_proofs_coalesced(ctx, cells.data(), proofs.data(), blob);
                if (rc) rcs[(size_t)t] = rc;
            }
        });
    for (auto& x : th) x.join();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int rc : rcs)
        if (rc) return rc;
    return RET_OK;
}

}  // extern "C"
