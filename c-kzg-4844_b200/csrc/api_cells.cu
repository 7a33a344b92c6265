// C ABI, EIP-7594 entry points (include/ckzg_b200.h): cells + FK20 proofs, batched over blobs.
#include <string.h>

#include "call.h"
#include "cells.h"

using namespace kzg;

namespace {

// monomial coefficients (device) -> 128 compressed proofs per blob at out48 (device)
int fk20_proofs_from_mono(Call& call, uint8_t* d_out48, const Fr* d_mono, uint64_t n) {
    Launch L = call.launch();
    uint32_t* d_S;
    G1 *d_u, *d_proofs;
    TRY(call.alloc(&d_S, n * 128 * 64 * 8));
    TRY(call.alloc(&d_u, n * 128));
    TRY(call.alloc(&d_proofs, n * 128));
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
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const uint64_t CHUNK = 512;  // bounds the per-blob scratch (0.6 MiB with proofs)
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

}  // extern "C"
