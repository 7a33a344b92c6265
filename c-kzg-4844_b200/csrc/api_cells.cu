// C ABI, EIP-7594 entry points (include/ckzg_b200.h): cells + FK20 proofs, batched over blobs.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <unordered_map>

#include "../src/host_sha256.h"
#include "call.h"
#include "cells.h"
#include "verify.h"

using namespace kzg;

namespace {

// monomial coefficients (device) -> 128 compressed proofs per blob at out48 (device).
// One stream: since the G1 FFTs work on quads of lanes and the MSMs add in affine coordinates, every phase
// fills the GPU at batch sizes of a few dozen blobs.  (Until r02 the batch was cut into up to four parts on
// side streams so that one part's MSMs ran under another's latency-bound FFT stages; measured again with the
// present kernels, 256 blobs: 1 part 36.3 ms, 2 parts 61.8 ms, 4 parts 39.8 ms.)
int fk20_proofs_from_mono(Call& call, uint8_t* d_out48, const Fr* d_mono, uint64_t n) {
    uint32_t* d_S;
    G1 *d_u, *d_proofs;
    TRY(call.alloc(&d_S, n * 128 * 64 * 8));
    TRY(call.alloc(&d_u, n * 128));
    TRY(call.alloc(&d_proofs, n * 128));
    Launch L = call.launch();
    TRY(launch_fk20_scalars(L, d_S, d_mono, n));
    TRY(launch_fk20_msm(L, d_u, d_S, n));
    TRY(launch_fk20_g1_ffts(L, d_proofs, d_u, n));
    TRY(launch_g1_compress(L, d_out48, d_proofs, n * 128));
    return RET_OK;
}

}  // namespace

namespace kzg {
int cells_fk20_proofs_from_mono(Call& call, uint8_t* d_out48, const Fr* d_mono, uint64_t n) { return fk20_proofs_from_mono(call, d_out48, d_mono, n); }
}  // namespace kzg

extern "C" {

int ckzg_b200_compute_cells_and_kzg_proofs_batch(ckzg_b200_ctx* ctx, uint8_t* cells, uint8_t* proofs, const uint8_t* blobs, uint64_t n, int mem, int* status) {
    if (!ctx || !blobs) return RET_BADARGS;
    if (!cells && !proofs) return RET_BADARGS;  // eip7594.c:72-74
    if (n == 0) return RET_OK;
    if (mem == CKZG_B200_HOST) {  // a context spanning devices: contiguous ranges of blobs, one per device
        Ctx* mc = reinterpret_cast<Ctx*>(ctx);
        const int parts = multi_parts(mc, n, 8);
        if (parts > 1)
            return multi_map(mc, n, parts, [&](ckzg_b200_ctx* dc, uint64_t f, uint64_t m) {
                return ckzg_b200_compute_cells_and_kzg_proofs_batch(dc, cells ? cells + f * 2 * BLOB_BYTES : nullptr, proofs ? proofs + f * 128 * 48 : nullptr,
                                                                    blobs + f * BLOB_BYTES, m, mem, status ? status + f : nullptr);
            });
    }
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    // bounds the per-blob scratch (0.6 MiB with proofs); big batches take 2048 at a time so that the one-thread-per-
    // butterfly G1 FFT stages (fk20_fft.cu, >= 1024 vectors) have 131 k chains to spread over the sub-partitions
    const uint64_t CHUNK = (proofs && n >= 2048) ? 2048 : 512;
    const uint64_t chunk = n < CHUNK ? n : CHUNK;
    const uint8_t* d_blobs;
    TRY(call.stage_in(&d_blobs, blobs, n * BLOB_BYTES, mem));
    int* d_bad;
    TRY(call.alloc(&d_bad, n));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, n * sizeof(int), call.stream));
    uint8_t *d_cells = nullptr, *d_proofs = nullptr;
    Fr* d_mono = nullptr;
    const bool dev = mem == CKZG_B200_DEVICE;
    if (cells) {
        if (dev)
            d_cells = cells;
        else
            TRY(call.alloc(&d_cells, chunk * 2 * BLOB_BYTES));
    }
    if (proofs) {
        TRY(call.alloc(&d_mono, chunk * N_BLOB));
        if (dev)
            d_proofs = proofs;
        else
            TRY(call.alloc(&d_proofs, chunk * 128 * 48));
    }
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t m = (n - off < chunk) ? n - off : chunk;
        uint8_t* c_out = d_cells ? (dev ? d_cells + off * 2 * BLOB_BYTES : d_cells) : nullptr;
        uint8_t* p_out = d_proofs ? (dev ? d_proofs + off * 128 * 48 : d_proofs) : nullptr;
        TRY(launch_blob_to_cells(L, c_out, d_mono, d_blobs + off * BLOB_BYTES, m, d_bad + off));
        if (proofs) TRY(fk20_proofs_from_mono(call, p_out, d_mono, m));
        if (!dev) {
            if (cells) KZG_CUDA_TRY(cudaMemcpyAsync(cells + off * 2 * BLOB_BYTES, c_out, m * 2 * BLOB_BYTES, cudaMemcpyDeviceToHost, call.stream));
            if (proofs) KZG_CUDA_TRY(cudaMemcpyAsync(proofs + off * 128 * 48, p_out, m * 128 * 48, cudaMemcpyDeviceToHost, call.stream));
        }
    }
    std::vector<int> bad(n);
    KZG_CUDA_TRY(cudaMemcpyAsync(bad.data(), d_bad, n * sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    int rc = RET_OK;
    for (uint64_t i = 0; i < n; i++) {
        int s = bad[i] ? RET_BADARGS : RET_OK;
        if (status) status[i] = s;
        if (s && !rc) rc = s;
    }
    return rc;
}


int ckzg_b200_recover_cells_and_kzg_proofs_batch(ckzg_b200_ctx* ctx, uint8_t* recovered_cells, uint8_t* recovered_proofs, const uint64_t* cell_indices, const uint8_t* cells,
                                                 uint64_t num_cells, uint64_t n, int mem, int* status) {
    if (!ctx || !recovered_cells || !cell_indices || !cells) return RET_BADARGS;
    if (n == 0) return RET_OK;
    // eip7594.c:191-213: count and index checks (host side, before anything else)
    if (num_cells > 128 || num_cells < 64) return RET_BADARGS;
    if (mem == CKZG_B200_HOST) {
        Ctx* mc = reinterpret_cast<Ctx*>(ctx);
        const int parts = multi_parts(mc, n, 8);
        if (parts > 1) {
            // index errors are BADARGS for the whole call before any work (eip7594.c:191-213): check all ranges first
            for (uint64_t b = 0; b < n; b++)
                for (uint64_t i = 0; i < num_cells; i++) {
                    const uint64_t v = cell_indices[b * num_cells + i];
                    if (v >= 128 || (i > 0 && v <= cell_indices[b * num_cells + i - 1])) return RET_BADARGS;
                }
            return multi_map(mc, n, parts, [&](ckzg_b200_ctx* dc, uint64_t f, uint64_t m) {
                return ckzg_b200_recover_cells_and_kzg_proofs_batch(dc, recovered_cells + f * 2 * BLOB_BYTES, recovered_proofs ? recovered_proofs + f * 128 * 48 : nullptr,
                                                                    cell_indices + f * num_cells, cells + f * num_cells * CELL_BYTES, num_cells, m, mem,
                                                                    status ? status + f : nullptr);
            });
        }
    }
    std::vector<int16_t> slot(n * 128, (int16_t)-1);
    std::vector<uint8_t> present(n * 128, 0);
    for (uint64_t b = 0; b < n; b++) {
        const uint64_t* idx = cell_indices + b * num_cells;
        for (uint64_t i = 0; i < num_cells; i++) {
            if (idx[i] >= 128) return RET_BADARGS;
            if (i > 0 && idx[i] <= idx[i - 1]) return RET_BADARGS;
            slot[b * 128 + idx[i]] = (int16_t)i;
            present[b * 128 + idx[i]] = 1;
        }
    }
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const bool dev = mem == CKZG_B200_DEVICE;
    const uint64_t CHUNK = (recovered_proofs && n >= 2048) ? 2048 : 256;
    const uint64_t chunk = n < CHUNK ? n : CHUNK;
    const uint8_t *d_cells_in, *d_slot, *d_present;
    TRY(call.stage_in(&d_cells_in, cells, n * num_cells * CELL_BYTES, mem));
    TRY(call.stage_in(&d_slot, (const uint8_t*)slot.data(), slot.size() * sizeof(int16_t), CKZG_B200_HOST));
    TRY(call.stage_in(&d_present, present.data(), present.size(), CKZG_B200_HOST));
    int* d_bad;
    TRY(call.alloc(&d_bad, n));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, n * sizeof(int), call.stream));
    uint8_t *d_cells_out, *d_proofs = nullptr, *scratch;
    Fr* d_mono = nullptr;
    if (dev)
        d_cells_out = recovered_cells;
    else
        TRY(call.alloc(&d_cells_out, chunk * 2 * BLOB_BYTES));
    TRY(call.alloc(&scratch, recover_scratch_bytes(chunk)));
    if (recovered_proofs) {
        TRY(call.alloc(&d_mono, chunk * N_BLOB));
        if (dev)
            d_proofs = recovered_proofs;
        else
            TRY(call.alloc(&d_proofs, chunk * 128 * 48));
    }
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t m = (n - off < chunk) ? n - off : chunk;
        uint8_t* c_out = dev ? d_cells_out + off * 2 * BLOB_BYTES : d_cells_out;
        uint8_t* p_out = d_proofs ? (dev ? d_proofs + off * 128 * 48 : d_proofs) : nullptr;
        TRY(launch_recover(L, c_out, d_mono, d_cells_in + off * num_cells * CELL_BYTES, (const int16_t*)d_slot + off * 128, d_present + off * 128, num_cells, m, d_bad + off,
                           scratch));
        if (recovered_proofs) TRY(fk20_proofs_from_mono(call, p_out, d_mono, m));
        if (!dev) {
            KZG_CUDA_TRY(cudaMemcpyAsync(recovered_cells + off * 2 * BLOB_BYTES, c_out, m * 2 * BLOB_BYTES, cudaMemcpyDeviceToHost, call.stream));
            if (recovered_proofs) KZG_CUDA_TRY(cudaMemcpyAsync(recovered_proofs + off * 128 * 48, p_out, m * 128 * 48, cudaMemcpyDeviceToHost, call.stream));
        }
    }
    std::vector<int> bad(n);
    KZG_CUDA_TRY(cudaMemcpyAsync(bad.data(), d_bad, n * sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    int rc = RET_OK;
    for (uint64_t i = 0; i < n; i++) {
        int s = bad[i] ? RET_BADARGS : RET_OK;
        if (status) status[i] = s;
        if (s && !rc) rc = s;
    }
    return rc;
}


int ckzg_b200_verify_cell_kzg_proof_batch(ckzg_b200_ctx* ctx, int* ok, const uint8_t* commitments, const uint64_t* cell_indices, const uint8_t* cells, const uint8_t* proofs,
                                          uint64_t n, int mem) {
    if (!ctx || !ok) return RET_BADARGS;
    *ok = 0;
    if (n == 0) {  // eip7594.c:852-855
        *ok = 1;
        return RET_OK;
    }
    if (!commitments || !cell_indices || !cells || !proofs) return RET_BADARGS;
    for (uint64_t i = 0; i < n; i++)
        if (cell_indices[i] >= 128) return RET_BADARGS;  // eip7594.c:861-864
    if (mem == CKZG_B200_HOST && !multi_inside_fanout()) {
        // Large batches are verified as independent SUB-BATCHES -- contiguous ranges of cells, each with its own
        // Fiat-Shamir challenge (each exactly the reference's verification of that range), verdicts AND-ed, any
        // invalid encoding -> BADARGS: a batch is valid iff every sub-batch is, so (return code, *ok) are the
        // reference's.  It is the reference's own parallel pattern (bindings/go/main_test.go:1037-1101), applied
        // inside the call: the one serial transcript hash (2112 B per cell, eip7594.c:405-474; 85 % of a 256-blob call
        // on one host core) becomes one hash per sub-batch on its own host thread, and the sub-batches spread over
        // the context's devices.  CKZG_B200_CELL_SUBBATCH=0 restores one challenge per call; =k sets the count per device.
        Ctx* mc = reinterpret_cast<Ctx*>(ctx);
        static const int sub_env = getenv("CKZG_B200_CELL_SUBBATCH") ? atoi(getenv("CKZG_B200_CELL_SUBBATCH")) : -1;
        const int per_dev = sub_env >= 0 ? sub_env : host_threads_default();
        const uint64_t want = (uint64_t)multi_device_count(mc) * (uint64_t)(per_dev < 1 ? 1 : per_dev);
        const uint64_t by_size = n / 4096;  // at least 32 blobs' worth of cells per sub-batch
        const uint64_t parts = sub_env == 0 ? (uint64_t)multi_parts(mc, n, 4096) : (by_size < want ? by_size : want);
        if (parts >= 2) {
            std::atomic<int> all_ok{1};
            int rc = multi_map(mc, n, (int)parts, [&](ckzg_b200_ctx* dc, uint64_t f, uint64_t m) {
                int o = 0;
                int r = ckzg_b200_verify_cell_kzg_proof_batch(dc, &o, commitments + 48 * f, cell_indices + f, cells + f * CELL_BYTES, proofs + 48 * f, m, mem);
                if (!o) all_ok.store(0);
                return r;
            });
            if (rc) return rc;
            *ok = all_ok.load();
            return RET_OK;
        }
    }
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const bool dev = mem == CKZG_B200_DEVICE;

    // host views of the byte inputs (the transcript is hashed on the host, src/host_sha256.c)
    std::vector<uint8_t> h_cm, h_cells, h_pf;
    const uint8_t *hc = commitments, *hcells = cells, *hp = proofs;
    if (dev) {
        h_cm.resize(n * 48);
        h_cells.resize(n * CELL_BYTES);
        h_pf.resize(n * 48);
        KZG_CUDA_TRY(cudaMemcpyAsync(h_cm.data(), commitments, n * 48, cudaMemcpyDeviceToHost, call.stream));
        KZG_CUDA_TRY(cudaMemcpyAsync(h_cells.data(), cells, n * CELL_BYTES, cudaMemcpyDeviceToHost, call.stream));
        KZG_CUDA_TRY(cudaMemcpyAsync(h_pf.data(), proofs, n * 48, cudaMemcpyDeviceToHost, call.stream));
        KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
        hc = h_cm.data();
        hcells = h_cells.data();
        hp = h_pf.data();
    }
    // deduplicate commitments in first-appearance order (eip7594.c:345-376)
    std::vector<uint8_t> uniq;
    std::vector<uint64_t> cm_index(n);
    {
        std::unordered_map<std::string, uint64_t> seen;
        seen.reserve(256);
        for (uint64_t i = 0; i < n; i++) {
            if (i > 0 && memcmp(hc + 48 * i, hc + 48 * (i - 1), 48) == 0) {  // row-major by blob: the usual case
                cm_index[i] = cm_index[i - 1];
                continue;
            }
            std::string key((const char*)hc + 48 * i, 48);
            auto it = seen.find(key);
            if (it == seen.end()) {
                uint64_t id = seen.size();
                seen.emplace(std::move(key), id);
                uniq.insert(uniq.end(), hc + 48 * i, hc + 48 * i + 48);
                cm_index[i] = id;
            } else {
                cm_index[i] = it->second;
            }
        }
    }
    const uint64_t u = uniq.size() / 48;

    // Validation does not need the challenge: it starts now and runs under the host's transcript hash.
    const uint8_t *d_cells, *d_pf, *d_uniq;
    TRY(call.stage_in(&d_pf, proofs, n * 48, mem));
    TRY(call.stage_in(&d_uniq, uniq.data(), uniq.size(), CKZG_B200_HOST));
    G1* d_table;
    int *d_flags;  // [0] bad input, [1] pairing verdict
    TRY(call.alloc(&d_table, verify_cells_table_points(n, u)));
    TRY(call.alloc(&d_flags, 2));
    KZG_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), call.stream));
    TRY(launch_verify_cells_validate(L, d_table, d_pf, n, d_uniq, u, d_flags));
    TRY(call.stage_in(&d_cells, cells, n * CELL_BYTES, mem));

    // challenge transcript (eip7594.c:390-482)
    uint8_t digest[32];
    {
        ckzg_host_sha256 h;
        ckzg_host_sha256_init(&h);
        uint8_t head[48] = {'R', 'C', 'K', 'Z', 'G', 'C', 'B', 'A', 'T', 'C', 'H', '_', '_', 'V', '1', '_'};
        const uint64_t vals[4] = {(uint64_t)N_BLOB, (uint64_t)CELL_FR, u, n};
        for (int v = 0; v < 4; v++)
            for (int i = 0; i < 8; i++) head[16 + 8 * v + i] = (uint8_t)(vals[v] >> (56 - 8 * i));
        ckzg_host_sha256_update(&h, head, 48);
        ckzg_host_sha256_update(&h, uniq.data(), uniq.size());
        for (uint64_t i = 0; i < n; i++) {
            uint8_t ids[16];
            for (int b = 0; b < 8; b++) {
                ids[b] = (uint8_t)(cm_index[i] >> (56 - 8 * b));
                ids[8 + b] = (uint8_t)(cell_indices[i] >> (56 - 8 * b));
            }
            ckzg_host_sha256_update(&h, ids, 16);
            ckzg_host_sha256_update(&h, hcells + i * CELL_BYTES, CELL_BYTES);
            ckzg_host_sha256_update(&h, hp + 48 * i, 48);
        }
        ckzg_host_sha256_final(&h, digest);
    }
    // CSR groupings: cells by column, cells by unique commitment
    std::vector<uint32_t> col_start(129, 0), col_items(n), cm_start(u + 1, 0), cm_items(n);
    std::vector<uint8_t> col_of(n);
    for (uint64_t i = 0; i < n; i++) {
        col_of[i] = (uint8_t)cell_indices[i];
        col_start[cell_indices[i] + 1]++;
        cm_start[cm_index[i] + 1]++;
    }
    for (int c = 0; c < 128; c++) col_start[c + 1] += col_start[c];
    for (uint64_t c = 0; c < u; c++) cm_start[c + 1] += cm_start[c];
    {
        std::vector<uint32_t> cc(col_start.begin(), col_start.end() - 1), mc(cm_start.begin(), cm_start.end() - 1);
        for (uint64_t i = 0; i < n; i++) {
            col_items[cc[cell_indices[i]]++] = (uint32_t)i;
            cm_items[mc[cm_index[i]]++] = (uint32_t)i;
        }
    }
    L.count(0, "transcript(host_sha)");  // what the host hash adds beyond the validation it runs beside

    const uint8_t *d_cs, *d_ci, *d_ms, *d_mi, *d_col;
    TRY(call.stage_in(&d_col, col_of.data(), col_of.size(), CKZG_B200_HOST));
    TRY(call.stage_in(&d_cs, (const uint8_t*)col_start.data(), col_start.size() * 4, CKZG_B200_HOST));
    TRY(call.stage_in(&d_ci, (const uint8_t*)col_items.data(), col_items.size() * 4, CKZG_B200_HOST));
    TRY(call.stage_in(&d_ms, (const uint8_t*)cm_start.data(), cm_start.size() * 4, CKZG_B200_HOST));
    TRY(call.stage_in(&d_mi, (const uint8_t*)cm_items.data(), cm_items.size() * 4, CKZG_B200_HOST));
    G1* d_AB;
    uint8_t* scratch;
    TRY(call.alloc(&d_AB, 2));
    TRY(call.alloc(&scratch, verify_cells_scratch_bytes(n, u)));
    TRY(launch_verify_cells(L, d_AB, d_table, d_cells, digest, d_col, (const uint32_t*)d_cs, (const uint32_t*)d_ci, (const uint32_t*)d_ms, (const uint32_t*)d_mi, n, u, d_flags,
                            scratch));
    // e(B, G2) == e(A, [tau^64]G2)   (eip7594.c:966); an invalid input leaves infinities behind, the
    // verdict is then ignored
    TRY(launch_pairing_check(L, d_flags + 1, d_AB + 0, d_AB + 1, nullptr, LINE_G2_TAU64, LINE_G2_GEN));
    int flags[2] = {0, 0};
    KZG_CUDA_TRY(cudaMemcpyAsync(flags, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    if (flags[0]) return RET_BADARGS;
    *ok = flags[1];
    return RET_OK;
}

}  // extern "C"
