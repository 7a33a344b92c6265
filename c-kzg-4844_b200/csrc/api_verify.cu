// C ABI, EIP-4844 proof and verification entry points (include/ckzg_b200.h), composed from the
// kernels of verify.cu / msm.cu / pairing.cu.  Host code here only sequences launches and copies.
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <functional>
#include <condition_variable>
#include <memory>
#include <thread>

#include "../src/host_sha256.h"
#include "call.h"
#include "verify.h"

using namespace kzg;

namespace kzg {
int verify_blob_batch_multi(Ctx* c, int* ok, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs, uint64_t n, int D);
}

namespace {

struct StatusOut {
    std::vector<int> host;
    int first = RET_OK;
};

// copy per-item flags back, OR them into the per-blob status
int collect_status(Call& call, const int* d_bad, uint64_t n, int* status) {
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

// shared tail of compute_kzg_proof / compute_blob_kzg_proof: z known -> y, quotient, MSM
int prove_batch(Call& call, uint8_t* proofs_out, uint8_t* ys_out, const uint8_t* d_blobs, const Fr* d_z, uint64_t n, int mem, int* d_bad) {
    Launch L = call.launch();
    const uint64_t CHUNK = 1024;  // bounds the 128 KiB/blob inverse + quotient scratch
    uint64_t chunk = n < CHUNK ? n : CHUNK;
    Fr *d_y, *d_inv;
    int* d_m;
    uint8_t *d_q, *d_zy, *d_proofs;
    TRY(call.alloc(&d_y, n));
    TRY(call.alloc(&d_zy, n * 64));
    TRY(call.alloc(&d_inv, chunk * N_BLOB));
    TRY(call.alloc(&d_m, chunk));
    TRY(call.alloc(&d_q, chunk * BLOB_BYTES));
    if (mem == CKZG_B200_DEVICE)
        d_proofs = proofs_out;
    else
        TRY(call.alloc(&d_proofs, n * 48));
    for (uint64_t off = 0; off < n; off += chunk) {
        uint64_t m = (n - off < chunk) ? n - off : chunk;
        const uint8_t* blobs = d_blobs + off * BLOB_BYTES;
        TRY(launch_evaluate(L, d_y + off, d_zy + off * 64, d_inv, d_m, blobs, d_z + off, m, d_bad + off, 1));
        TRY(launch_quotient(L, d_q, blobs, d_z + off, d_y + off, d_inv, d_m, m));
        TRY(commit_scalars_batch(call, d_proofs + off * 48, d_q, false, m, nullptr));
    }
    if (mem != CKZG_B200_DEVICE) KZG_CUDA_TRY(cudaMemcpyAsync(proofs_out, d_proofs, n * 48, cudaMemcpyDeviceToHost, call.stream));
    if (ys_out) {
        // y bytes sit at zy[i*64+32 ..): gather with a strided 2D copy
        cudaMemcpyKind kind = (mem == CKZG_B200_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        KZG_CUDA_TRY(cudaMemcpy2DAsync(ys_out, 32, d_zy + 32, 64, 32, n, kind, call.stream));
    }
    return RET_OK;
}

// Per-blob stage of the verifier: validate points, challenges z, evaluations y.
struct Stage1 {
    G1Affine *cm = nullptr, *pf = nullptr;  // pf = pts, cm = pts + n (one array: vmsm.cu point layout)
    Fr *z = nullptr, *y = nullptr;
    uint8_t* zy = nullptr;
    int* bad = nullptr;  // single flag
    // table of shifted points for the bucket form of the linear combination (vmsm.cu), written by the
    // validation kernels as a by-product of the subgroup test
    bool want_shift = false;
    G1* table = nullptr;
    // Streamed z||y (batch verification): when h_zy (pinned, n x 64) is set, the evaluations run in chunks and every
    // chunk's z||y is copied to h_zy on a side stream as soon as it exists; zy_done[k] fires when chunk k (blobs
    // [k * zy_chunk, ...)) has landed, so the host hashes the batch transcript BEHIND the GPU instead of after it.
    uint8_t* h_zy = nullptr;
    std::vector<std::pair<uint64_t, uint64_t>> zy_range;  // (first blob, count) of zy_done[k]
    std::vector<cudaEvent_t> zy_done;
    ~Stage1() {
        for (cudaEvent_t e : zy_done)
            if (e) cudaEventDestroy(e);
    }
};
// CKZG_B200_RLC=points forces the one-multiplication-per-point linear combination (A/B comparison)
bool rlc_use_vmsm() {
    static const bool on = !(getenv("CKZG_B200_RLC") && strcmp(getenv("CKZG_B200_RLC"), "points") == 0);
    return on;
}
// `blobs` lives in `mem` space.  The per-blob Fiat-Shamir hash is latency bound (2050 dependent SHA-256
// blocks per thread: ~4 ms whatever the batch size, on 1 warp per SM), so it runs on side streams
// concurrently with the point validations; HOST blobs are uploaded in chunks on a copy stream and every
// chunk's (hash -> evaluate) chain starts as soon as its bytes have landed (copy engine || SMs).
// With per-kernel profiling on (level 2) everything runs on the call's stream so event attribution is exact.
// after the evaluation of blobs [off, off + m) has been enqueued on the call stream: copy their z||y to the host on `cpz`
static int stage1_stream_zy(Call& call, Stage1& s, cudaStream_t cpz, uint64_t off, uint64_t m) {
    cudaEvent_t ready = nullptr, done = nullptr;
    KZG_CUDA_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaEventRecord(ready, call.stream);
    cudaStreamWaitEvent(cpz, ready, 0);
    cudaEventDestroy(ready);
    KZG_CUDA_TRY(cudaMemcpyAsync(s.h_zy + off * 64, s.zy + off * 64, m * 64, cudaMemcpyDeviceToHost, cpz));
    KZG_CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    cudaEventRecord(done, cpz);
    s.zy_done.push_back(done);
    s.zy_range.push_back({off, m});
    return RET_OK;
}

int verify_stage1(Call& call, Stage1& s, const uint8_t* blobs, const uint8_t* d_cm, const uint8_t* d_pf, uint64_t n, int mem) {
    Launch L = call.launch();
    const bool stream_zy = s.h_zy != nullptr && !call.trace_kernels && n >= 1024;
    cudaStream_t cpz = nullptr;
    if (stream_zy && !(cpz = call.fork())) return RET_ERROR;
    const bool host = (mem == CKZG_B200_HOST);
    uint8_t* d_up = nullptr;
    if (host) TRY(call.alloc(&d_up, n * BLOB_BYTES));
    const uint8_t* d_blobs = host ? d_up : blobs;
    TRY(call.alloc(&s.pf, 2 * n + 1));
    s.cm = s.pf + n;
    if (s.want_shift) {
        TRY(call.alloc(&s.table, vmsm_table_points(n)));
        TRY(vmsm_place_generator(L, s.table, n));
    }
    TRY(call.alloc(&s.z, n));
    TRY(call.alloc(&s.y, n));
    TRY(call.alloc(&s.zy, n * 64));
    TRY(call.alloc(&s.bad, 1));
    KZG_CUDA_TRY(cudaMemsetAsync(s.bad, 0, sizeof(int), call.stream));
    if (host && (call.trace_kernels || n < 64)) {  // serial form
        if (host) TRY(call.upload(d_up, blobs, n * BLOB_BYTES, call.stream));
        if (s.want_shift)
            TRY(launch_g1_validate2_levels(L, s.cm, d_cm, s.pf, d_pf, n, s.bad, s.table));
        else
            TRY(launch_g1_validate2(L, s.cm, d_cm, s.pf, d_pf, n, s.bad));
        TRY(launch_blob_challenges(L, s.z, s.zy, d_blobs, d_cm, n));
        TRY(launch_evaluate(L, s.y, s.zy, nullptr, nullptr, d_blobs, s.z, n, s.bad, 0));
        return RET_OK;
    }
    static const int stage1_mode = getenv("CKZG_B200_STAGE1") ? atoi(getenv("CKZG_B200_STAGE1")) : 1;
    if (!host && stage1_mode == 1) {
        // Device-resident blobs: hash and validation share one kernel (one warp per sub-partition), the
        // evaluations follow on the same stream.
        TRY(launch_stage1_fused(L, s.z, s.zy, d_blobs, s.cm, d_cm, s.pf, d_pf, n, s.bad, s.table));
        call.mark_on(call.stream, "stage:t_hash_done");
        int rc = RET_OK;
        if (stream_zy) {
            const uint64_t zy_chunk = (n + 3) / 4;  // four chunks: the host hashes chunk k while the GPU evaluates chunk k + 1
            for (uint64_t off = 0; off < n && rc == RET_OK; off += zy_chunk) {
                const uint64_t m = (n - off < zy_chunk) ? n - off : zy_chunk;
                rc = launch_evaluate(L, s.y + off, s.zy + off * 64, nullptr, nullptr, d_blobs + off * BLOB_BYTES, s.z + off, m, s.bad, 0);
                if (rc == RET_OK) rc = stage1_stream_zy(call, s, cpz, off, m);
            }
        } else {
            rc = launch_evaluate(L, s.y, s.zy, nullptr, nullptr, d_blobs, s.z, n, s.bad, 0);
        }
        call.mark_on(call.stream, "stage:t_evaluate_done");
        return rc;
    }
    // fork: side streams start after the allocations / memset enqueued so far.
    // Measured on B200 (tools/gpu_probe.py modes, n = 4096 device-resident): everything on one stream
    // 16.4 ms; hash -> evaluate chains on a side stream next to the validations 17.6 ms (the two
    // throughput kernels get in each other's way); ONLY the latency-bound hashes on side streams, the
    // validations and then the evaluations on the main stream: 14.3 ms.  That is the arrangement here.
    const uint64_t CH = host ? 512 : n;
    // Segments of the batch.  Ordinary segments are chunks of CH blobs: one copy, one hash launch when it has landed.
    // The TAIL of a pinned host batch -- the last 1024 blobs, whose bytes take 2.4 ms to arrive, as long as one blob
    // takes to hash (2050 dependent SHA-256 blocks) -- travels in column pieces instead: piece p = bytes
    // [16 KiB p, 16 KiB (p + 1)) of EVERY tail blob (one strided copy), hashed piece by piece behind the copies with the
    // SHA state carried between the launches (verify.cu launch_blob_challenges_range).  When the batch's last byte
    // lands, 258 of the 2050 blocks of each tail blob are left to hash (0.3 ms) instead of a whole 2.1 ms chain; earlier
    // chunks finish their hashes under the copies that follow them anyway.  16.4 -> 14.9 ms per 4096-blob call
    // (tools/e2e_probe.py, profiles/e2e_probe_R2y.log).
    // This arrangement (tail pieces, four hash streams) is used when the process runs with at most four hardware work
    // queues (CUDA_DEVICE_MAX_CONNECTIONS <= 4: the library's default when it is loaded before CUDA is initialised,
    // api.cu); with more queues it -- like any arrangement but the one of round 1 -- ran into events firing ~11 ms late
    // (profiles/e2e_probe_R2k ... R2x.log), so there the round-1 arrangement stays: one hash stream per chunk, no pieces.
    // CKZG_B200_TAIL_PIECES / CKZG_B200_TAIL_BLOBS / CKZG_B200_HASH_STREAMS override.
    struct Seg {
        uint64_t off, m;
        bool pieces;
    };
    std::vector<Seg> segs;
    static const int hw_queues = (getenv("CUDA_DEVICE_MAX_CONNECTIONS") && atoi(getenv("CUDA_DEVICE_MAX_CONNECTIONS")) >= 1) ? atoi(getenv("CUDA_DEVICE_MAX_CONNECTIONS")) : 8;
    static const int tail_pieces = getenv("CKZG_B200_TAIL_PIECES") ? atoi(getenv("CKZG_B200_TAIL_PIECES")) : (hw_queues <= 4 ? 8 : 0);
    static const uint64_t tail_blobs = getenv("CKZG_B200_TAIL_BLOBS") ? (uint64_t)atoll(getenv("CKZG_B200_TAIL_BLOBS")) : 1024;
    uint64_t tail_start = n;
    if (host && tail_pieces >= 2 && tail_pieces <= 64 && (N_BLOB * 32 / 64) % tail_pieces == 0 && tail_blobs >= 128) {
        const uint64_t want = n < tail_blobs ? n : tail_blobs;
        const uint64_t start = ((n - want) / CH) * CH;  // on a chunk boundary
        if ((n - start) * BLOB_BYTES >= (16u << 20) && !host_ptr_is_pageable(blobs + start * BLOB_BYTES)) tail_start = start;
    }
    for (uint64_t off = 0; off < tail_start; off += CH) segs.push_back({off, (tail_start - off < CH) ? tail_start - off : CH, false});
    if (tail_start < n) segs.push_back({tail_start, n - tail_start, true});
    const int nsegs = (int)segs.size();
    static const int max_side = (getenv("CKZG_B200_HASH_STREAMS") && atoi(getenv("CKZG_B200_HASH_STREAMS")) >= 1 && atoi(getenv("CKZG_B200_HASH_STREAMS")) <= 8) ? atoi(getenv("CKZG_B200_HASH_STREAMS")) : (hw_queues <= 4 ? 4 : 8);
    const int nside = std::min(max_side, nsegs);
    // side streams are forked from (ordered after) the call stream and joined / destroyed by the Call on
    // every exit path, so no early return below can leave a kernel reading released scratch
    cudaStream_t side[8], copy = nullptr, tail_stream = nullptr;
    int rc = RET_OK;
    for (int i = 0; i < nside; i++)
        if (!(side[i] = call.fork())) return RET_ERROR;
    if (host && !(copy = call.fork())) return RET_ERROR;
    if (tail_start < n && nsegs > 1 && !(tail_stream = call.fork())) return RET_ERROR;
    std::vector<cudaEvent_t> hashed(nsegs, nullptr);
    uint32_t* d_states = nullptr;
    if (tail_start < n) TRY(call.alloc(&d_states, (n - tail_start) * 8));
    // EXPERIMENT, OFF (CKZG_B200_VALIDATE_FIRST=1 enables it): the point validation enqueued on the main stream BEFORE the
    // segments.  Enqueued after them (the default, below) it starts 5.2 ms into the call: streams share hardware work
    // queues, a queue hands out its entries in order, and the validation sits behind another stream's wait for a chunk
    // that has not landed yet (device timeline, profiles/e2e_probe_R3a.log); the evaluations of the first four chunks
    // then run in one burst beside the last chunks' hashes.  Moving it to the front did start it at once -- and the
    // hash of chunk 4 then started 15.5 ms into the call instead of 6.4 (profiles/e2e_probe_R3b.log): 22.1 ms per call
    // against 14.4.  Which stream lands behind which in a queue is not under the library's control; the order below is
    // the one whose measured timeline is good.
    // (2 = a third arrangement: the validation on a side stream of its own, forked here, joined before the first evaluation)
    static const int validate_mode = getenv("CKZG_B200_VALIDATE_FIRST") ? atoi(getenv("CKZG_B200_VALIDATE_FIRST")) : 0;
    const bool validate_first = validate_mode == 1 || validate_mode == 2;
    cudaEvent_t validated = nullptr;
    if (validate_mode == 1) {
        TRY(s.want_shift ? launch_g1_validate2_levels(L, s.cm, d_cm, s.pf, d_pf, n, s.bad, s.table) : launch_g1_validate2(L, s.cm, d_cm, s.pf, d_pf, n, s.bad));
        call.mark_on(call.stream, "stage:t_validate_done");
    } else if (validate_mode == 2) {
        cudaStream_t vs = call.fork();
        if (!vs) return RET_ERROR;
        Launch Lv = call.launch_on(vs);
        TRY(s.want_shift ? launch_g1_validate2_levels(Lv, s.cm, d_cm, s.pf, d_pf, n, s.bad, s.table) : launch_g1_validate2(Lv, s.cm, d_cm, s.pf, d_pf, n, s.bad));
        KZG_CUDA_TRY(cudaEventCreateWithFlags(&validated, cudaEventDisableTiming));
        cudaEventRecord(validated, vs);
        call.mark_on(vs, "stage:t_validate_done");
    }
    for (int c = 0; c < nsegs && rc == RET_OK; c++) {
        const uint64_t off = segs[c].off, m = segs[c].m;
        cudaStream_t st = (segs[c].pieces && tail_stream) ? tail_stream : side[c % nside];
        Launch Ls = call.launch_on(st);
        if (segs[c].pieces) {
            const int P = tail_pieces;
            const size_t piece_bytes = BLOB_BYTES / P;  // 16 KiB at P = 8
            const int blocks_per_piece = (int)(piece_bytes / 64);
            for (int p = 0; p < P && rc == RET_OK; p++) {
                cudaEvent_t landed;
                if (cudaMemcpy2DAsync(d_up + off * BLOB_BYTES + p * piece_bytes, BLOB_BYTES, blobs + off * BLOB_BYTES + p * piece_bytes, BLOB_BYTES, piece_bytes, m,
                                      cudaMemcpyHostToDevice, copy) != cudaSuccess ||
                    cudaEventCreateWithFlags(&landed, cudaEventDisableTiming) != cudaSuccess) {
                    rc = RET_ERROR;
                    break;
                }
                cudaEventRecord(landed, copy);
                cudaStreamWaitEvent(st, landed, 0);
                cudaEventDestroy(landed);
                // block k >= 1 needs blob bytes up to 64 k + 32: bytes [0, (p + 1) * piece) allow blocks [.., (p + 1) * blocks_per_piece)
                const int k0 = p * blocks_per_piece, k1 = (p == P - 1) ? blob_challenge_blocks() : (p + 1) * blocks_per_piece;
                rc = launch_blob_challenges_range(Ls, s.z + off, s.zy + off * 64, d_blobs + off * BLOB_BYTES, d_cm + off * 48, m, k0, k1, d_states);
            }
        } else {
            if (host) {
                cudaEvent_t landed;
                // pinned sources: one DMA per chunk; pageable sources: staged through the pinned ring by the host threads
                if (call.upload(d_up + off * BLOB_BYTES, blobs + off * BLOB_BYTES, m * BLOB_BYTES, copy) != RET_OK ||
                    cudaEventCreateWithFlags(&landed, cudaEventDisableTiming) != cudaSuccess) {
                    rc = RET_ERROR;
                    break;
                }
                cudaEventRecord(landed, copy);
                cudaStreamWaitEvent(st, landed, 0);
                cudaEventDestroy(landed);
                if (c == 0) call.mark_on(copy, "stage:t_copy0_done");
                if (c == 0) call.mark_on(st, "stage:t_hash0_start");
                if (c == 1) call.mark_on(st, "stage:t_hash1_start");
            }
            rc = launch_blob_challenges(Ls, s.z + off, s.zy + off * 64, d_blobs + off * BLOB_BYTES, d_cm + off * 48, m);
        }
        if (rc) break;
        if (cudaEventCreateWithFlags(&hashed[c], cudaEventDisableTiming) != cudaSuccess) {
            rc = RET_ERROR;
            break;
        }
        cudaEventRecord(hashed[c], st);
        if (c == 0) call.mark_on(st, "stage:t_first_chunk_hashed");
        if (c == nsegs - 1) call.mark_on(st, "stage:t_hash_done");
    }
    if (copy) call.mark_on(copy, "stage:t_upload_done");
    // main stream: point validation (independent of the blobs), then each segment's evaluation as soon as
    // its challenges exist
    if (rc == RET_OK && !validate_first) {
        rc = s.want_shift ? launch_g1_validate2_levels(L, s.cm, d_cm, s.pf, d_pf, n, s.bad, s.table) : launch_g1_validate2(L, s.cm, d_cm, s.pf, d_pf, n, s.bad);
        call.mark_on(call.stream, "stage:t_validate_done");
    }
    if (validated) {
        cudaStreamWaitEvent(call.stream, validated, 0);
        cudaEventDestroy(validated);
    }
    call.host_mark("host:t_stage1_enqueued");
    for (int c = 0; c < nsegs; c++) {
        const uint64_t off = segs[c].off, m = segs[c].m;
        if (!hashed[c]) continue;
        if (c == nsegs - 1) call.mark_on(call.stream, "stage:t_earlier_evaluations_done");
        cudaStreamWaitEvent(call.stream, hashed[c], 0);
        cudaEventDestroy(hashed[c]);
        if (c == nsegs - 1) call.mark_on(call.stream, "stage:t_last_evaluation_start");
        if (rc == RET_OK) rc = launch_evaluate(L, s.y + off, s.zy + off * 64, nullptr, nullptr, d_blobs + off * BLOB_BYTES, s.z + off, m, s.bad, 0);
        if (c == nsegs - 1) call.mark_on(call.stream, "stage:t_last_evaluation_done");
        if (rc == RET_OK && stream_zy) rc = stage1_stream_zy(call, s, cpz, off, m);
    }
    return rc;
}

// r = hash_to_bls_field(SHA256("RCKZGBATCH___V1_" || u64be(4096) || u64be(n) || n x (C || z || y || proof)))
// (compute_r_powers_for_verify_kzg_proof_batch, eip4844.c:612-668).  The transcript is one serial hash
// chain: hashed on the host (src/host_sha256.c explains why); the 32-byte digest goes back to the
// device, which reduces it mod r.  Inputs are host pointers; zy is n x 64 (z || y).
struct TranscriptHasher {
    ckzg_host_sha256 h;
    void begin(uint64_t n) {
        ckzg_host_sha256_init(&h);
        uint8_t head[32] = {'R', 'C', 'K', 'Z', 'G', 'B', 'A', 'T', 'C', 'H', '_', '_', '_', 'V', '1', '_'};
        for (int i = 0; i < 8; i++) {
            head[16 + i] = (uint8_t)((uint64_t)N_BLOB >> (56 - 8 * i));
            head[24 + i] = (uint8_t)(n >> (56 - 8 * i));
        }
        ckzg_host_sha256_update(&h, head, 32);
    }
    void feed(const uint8_t* cm, const uint8_t* zy, const uint8_t* pf, uint64_t first, uint64_t count) {
        for (uint64_t i = first; i < first + count; i++) {
            ckzg_host_sha256_update(&h, cm + 48 * i, 48);
            ckzg_host_sha256_update(&h, zy + 64 * i, 64);
            ckzg_host_sha256_update(&h, pf + 48 * i, 48);
        }
    }
    void finish(uint8_t digest[32]) { ckzg_host_sha256_final(&h, digest); }
};
void transcript_digest(uint8_t digest[32], const uint8_t* cm, const uint8_t* zy, const uint8_t* pf, const uint8_t* tuples, uint64_t n) {
    ckzg_host_sha256 h;
    ckzg_host_sha256_init(&h);
    uint8_t head[32] = {'R', 'C', 'K', 'Z', 'G', 'B', 'A', 'T', 'C', 'H', '_', '_', '_', 'V', '1', '_'};
    for (int i = 0; i < 8; i++) {
        head[16 + i] = (uint8_t)((uint64_t)N_BLOB >> (56 - 8 * i));
        head[24 + i] = (uint8_t)(n >> (56 - 8 * i));
    }
    ckzg_host_sha256_update(&h, head, 32);
    if (tuples) {
        ckzg_host_sha256_update(&h, tuples, 160 * n);
    } else {
        for (uint64_t i = 0; i < n; i++) {
            ckzg_host_sha256_update(&h, cm + 48 * i, 48);
            ckzg_host_sha256_update(&h, zy + 64 * i, 64);
            ckzg_host_sha256_update(&h, pf + 48 * i, 48);
        }
    }
    ckzg_host_sha256_final(&h, digest);
}
int read_flag(Call& call, const int* d_flag, int* out) {
    KZG_CUDA_TRY(cudaMemcpyAsync(out, d_flag, sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    return RET_OK;
}

}  // namespace

extern "C" {

int ckzg_b200_compute_kzg_proof_batch(ckzg_b200_ctx* ctx, uint8_t* proofs, uint8_t* ys, const uint8_t* blobs, const uint8_t* zs, uint64_t n, int mem, int* status) {
    if (!ctx || !proofs || !ys || !blobs || !zs) return RET_BADARGS;
    if (n == 0) return RET_OK;
    if (mem == CKZG_B200_HOST) {
        Ctx* mc = reinterpret_cast<Ctx*>(ctx);
        const int parts = multi_parts(mc, n, 16);
        if (parts > 1)
            return multi_map(mc, n, parts, [&](ckzg_b200_ctx* dc, uint64_t f, uint64_t m) {
                return ckzg_b200_compute_kzg_proof_batch(dc, proofs + 48 * f, ys + 32 * f, blobs + f * BLOB_BYTES, zs + 32 * f, m, mem, status ? status + f : nullptr);
            });
    }
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const uint8_t *d_blobs, *d_zs;
    TRY(call.stage_in(&d_blobs, blobs, n * BLOB_BYTES, mem));
    TRY(call.stage_in(&d_zs, zs, n * 32, mem));
    int* d_bad;
    Fr* d_z;
    TRY(call.alloc(&d_bad, n));
    TRY(call.alloc(&d_z, n));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, n * sizeof(int), call.stream));
    TRY(launch_z_from_bytes(L, d_z, nullptr, d_zs, n, d_bad));
    TRY(prove_batch(call, proofs, ys, d_blobs, d_z, n, mem, d_bad));
    return collect_status(call, d_bad, n, status);
}

int ckzg_b200_compute_blob_kzg_proof_batch(ckzg_b200_ctx* ctx, uint8_t* proofs, const uint8_t* blobs, const uint8_t* commitments, uint64_t n, int mem, int* status) {
    if (!ctx || !proofs || !blobs || !commitments) return RET_BADARGS;
    if (n == 0) return RET_OK;
    if (mem == CKZG_B200_HOST) {
        Ctx* mc = reinterpret_cast<Ctx*>(ctx);
        const int parts = multi_parts(mc, n, 16);
        if (parts > 1)
            return multi_map(mc, n, parts, [&](ckzg_b200_ctx* dc, uint64_t f, uint64_t m) {
                return ckzg_b200_compute_blob_kzg_proof_batch(dc, proofs + 48 * f, blobs + f * BLOB_BYTES, commitments + 48 * f, m, mem, status ? status + f : nullptr);
            });
    }
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const uint8_t *d_blobs, *d_cm;
    TRY(call.stage_in(&d_blobs, blobs, n * BLOB_BYTES, mem));
    TRY(call.stage_in(&d_cm, commitments, n * 48, mem));
    int* d_bad;
    Fr* d_z;
    G1Affine* d_cm_pts;
    uint8_t* d_zy;
    TRY(call.alloc(&d_bad, n));
    TRY(call.alloc(&d_z, n));
    TRY(call.alloc(&d_cm_pts, n));
    TRY(call.alloc(&d_zy, n * 64));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, n * sizeof(int), call.stream));
    TRY(launch_g1_validate(L, d_cm_pts, d_cm, n, d_bad, 1));  // bytes_to_kzg_commitment, eip4844.c:520
    TRY(launch_blob_challenges(L, d_z, d_zy, d_blobs, d_cm, n));
    TRY(prove_batch(call, proofs, nullptr, d_blobs, d_z, n, mem, d_bad));
    return collect_status(call, d_bad, n, status);
}

// context-free calls (hash / encoding helpers of the reference that take no KZGSettings) run on a bare
// context bound to the current device
// (one immutable bare context per device, created once under a lock: concurrent callers with different current
// devices never see each other's device -- the reference's helpers are re-entrant and so are these)
static Ctx* bare_ctx() {
    static std::mutex mu;
    static Ctx* table[64] = {nullptr};
    int dev = 0;
    const char* env = getenv("CKZG_B200_DEVICE");
    if (env)
        dev = atoi(env);
    else if (cudaGetDevice(&dev) != cudaSuccess)
        dev = 0;
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> g(mu);
    if (!table[dev]) {
        table[dev] = new Ctx();
        table[dev]->device = dev;
    }
    return table[dev];
}

int ckzg_b200_hash_to_bls_field(uint8_t* out32, const uint8_t* digest32) {
    if (!out32 || !digest32) return RET_BADARGS;
    Call call(bare_ctx());
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const uint8_t* d_digest;
    TRY(call.stage_in(&d_digest, digest32, 32, CKZG_B200_HOST));
    Fr* d_r;
    uint8_t* d_out;
    TRY(call.alloc(&d_r, 1));
    TRY(call.alloc(&d_out, 64));
    TRY(launch_r_from_digest(L, d_r, d_digest));
    TRY(launch_fr_to_bytes(L, d_out, d_r, 1));
    KZG_CUDA_TRY(cudaMemcpyAsync(out32, d_out, 32, cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    return RET_OK;
}

int ckzg_b200_validate_g1(int* ok, const uint8_t* p48) {
    if (!ok || !p48) return RET_BADARGS;
    *ok = 0;
    Call call(bare_ctx());
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const uint8_t* d_in;
    TRY(call.stage_in(&d_in, p48, 48, CKZG_B200_HOST));
    G1Affine* d_pt;
    int* d_bad;
    TRY(call.alloc(&d_pt, 1));
    TRY(call.alloc(&d_bad, 1));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), call.stream));
    TRY(launch_g1_validate(L, d_pt, d_in, 1, d_bad, 0));
    int bad = 0;
    KZG_CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    *ok = bad ? 0 : 1;
    return RET_OK;
}

int ckzg_b200_compute_challenge(ckzg_b200_ctx* ctx, uint8_t* out32, const uint8_t* blob, const uint8_t* commitment48) {
    if (!out32 || !blob || !commitment48) return RET_BADARGS;
    Call call(ctx ? reinterpret_cast<Ctx*>(ctx) : bare_ctx());
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const uint8_t *d_blob, *d_cm;
    TRY(call.stage_in(&d_blob, blob, BLOB_BYTES, CKZG_B200_HOST));
    TRY(call.stage_in(&d_cm, commitment48, 48, CKZG_B200_HOST));
    Fr* d_z;
    uint8_t* d_zy;
    TRY(call.alloc(&d_z, 1));
    TRY(call.alloc(&d_zy, 64));
    TRY(launch_blob_challenges(L, d_z, d_zy, d_blob, d_cm, 1));
    KZG_CUDA_TRY(cudaMemcpyAsync(out32, d_zy, 32, cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    return RET_OK;
}

int ckzg_b200_verify_blob_kzg_proof_batch(ckzg_b200_ctx* ctx, int* ok, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs, uint64_t n, int mem) {
    if (!ctx || !ok) return RET_BADARGS;
    *ok = 0;
    if (n == 0) {  // eip4844.c:791
        *ok = 1;
        return RET_OK;
    }
    if (!blobs || !commitments || !proofs) return RET_BADARGS;
    {
        // a context that spans devices (CKZG_B200_DEVICES) shards HOST batches: >= 256 blobs per device
        Ctx* c = reinterpret_cast<Ctx*>(ctx);
        const uint64_t D = multi_inside_fanout() ? 1 : std::min<uint64_t>(c->peers.size(), n / 256);
        if (mem == CKZG_B200_HOST && D >= 2) return verify_blob_batch_multi(c, ok, blobs, commitments, proofs, n, (int)D);
    }
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    const uint8_t *d_blobs, *d_cm, *d_pf;
    TRY(call.stage_in(&d_cm, commitments, n * 48, mem));
    TRY(call.stage_in(&d_pf, proofs, n * 48, mem));
    Stage1 s;
    (void)d_blobs;
    bool use_r = n > 1;  // n == 1: the single-proof equation, no challenge (eip4844.c:798)
    s.want_shift = use_r && rlc_use_vmsm();
    // Host copies for the batch transcript (pinned): z||y and the flag after stage 1; for device-resident
    // inputs also the commitments and proofs, fetched on a side stream while stage 1 runs.
    uint8_t* pin = nullptr;
    cudaEvent_t fetched = nullptr;
    TRY(call.pin(&pin, n * 160 + 64));
    uint8_t *h_zy = pin, *h_c = pin + n * 64, *h_p = pin + n * 112;
    int* h_bad = (int*)(pin + n * 160);
    if (use_r && mem == CKZG_B200_DEVICE) {
        cudaStream_t cp = call.fork();  // ordered after the caller's producer through the call stream
        if (!cp || cudaEventCreateWithFlags(&fetched, cudaEventDisableTiming) != cudaSuccess) return RET_ERROR;
        cudaMemcpyAsync(h_c, d_cm, n * 48, cudaMemcpyDeviceToHost, cp);
        cudaMemcpyAsync(h_p, d_pf, n * 48, cudaMemcpyDeviceToHost, cp);
        cudaEventRecord(fetched, cp);
    }
    if (use_r) s.h_zy = h_zy;  // stream z||y to the host chunk by chunk (n >= 1024), hashed behind the evaluations
    const uint8_t* hc = (mem == CKZG_B200_DEVICE) ? h_c : commitments;
    const uint8_t* hp = (mem == CKZG_B200_DEVICE) ? h_p : proofs;
    TranscriptHasher th;
    th.begin(n);
    size_t fed = 0;
    int rc1 = verify_stage1(call, s, blobs, d_cm, d_pf, n, mem);
    if (rc1) {
        if (fetched) {
            cudaEventSynchronize(fetched);
            cudaEventDestroy(fetched);
        }
        return rc1;
    }
    const bool stage_marks = call.profiling && !call.trace_kernels;  // level 1: stage boundaries of the concurrent form
    if (stage_marks) call.mark("stage:per_blob(validate|hash,evaluate)");
    // the reference stops at the first invalid input (eip4844.c:813-831) before any pairing work
    const bool streamed = !s.zy_done.empty();
    if (use_r && !streamed) cudaMemcpyAsync(h_zy, s.zy, n * 64, cudaMemcpyDeviceToHost, call.stream);
    cudaMemcpyAsync(h_bad, s.bad, sizeof(int), cudaMemcpyDeviceToHost, call.stream);
    uint8_t digest[32] = {0};
    cudaError_t se = cudaSuccess;
    if (fetched) {
        se = cudaEventSynchronize(fetched);
        cudaEventDestroy(fetched);
    }
    if (streamed && se == cudaSuccess) {
        // the rest of the batch transcript (eip4844.c:612-668): chunks not yet fed while stage 1 was being driven
        for (; fed < s.zy_done.size() && se == cudaSuccess; fed++) {
            se = cudaEventSynchronize(s.zy_done[fed]);
            if (se == cudaSuccess) th.feed(hc, h_zy, hp, s.zy_range[fed].first, s.zy_range[fed].second);
        }
        th.finish(digest);
    }
    if (se == cudaSuccess) se = cudaStreamSynchronize(call.stream);
    KZG_CUDA_TRY(se);
    call.host_mark("host:t_stage1_synced");
    if (*h_bad) return RET_BADARGS;

    Fr* d_r;
    TRY(call.alloc(&d_r, 1));
    if (use_r) {
        if (!streamed) transcript_digest(digest, hc, h_zy, hp, nullptr, n);
        L.count(0, "transcript(d2h,host_sha)");
        if (!s.want_shift) {
            uint8_t* d_digest;
            TRY(call.alloc(&d_digest, 32));
            memcpy(pin + n * 160 + 16, digest, 32);  // pinned: no synchronisation needed before the launch
            KZG_CUDA_TRY(cudaMemcpyAsync(d_digest, pin + n * 160 + 16, 32, cudaMemcpyHostToDevice, call.stream));
            TRY(launch_r_from_digest(L, d_r, d_digest));
        }
    }
    if (stage_marks) call.mark("stage:transcript(d2h,host_sha,r)");
    G1* d_AB;
    void* scratch;
    int* d_ok;
    TRY(call.alloc(&d_AB, 2));
    TRY(call.alloc((uint8_t**)&scratch, s.want_shift ? rlc_vmsm_scratch_bytes(n) : rlc_scratch_bytes(n)));
    TRY(call.alloc(&d_ok, 1));
    if (s.want_shift)
        TRY(launch_rlc_vmsm(L, d_AB, s.table, s.z, s.y, digest, 0, n, scratch));  // the digest travels as a kernel argument
    else
        TRY(launch_rlc(L, d_AB, s.cm, s.pf, s.z, s.y, d_r, use_r, 0, n, scratch));
    if (stage_marks) call.mark("stage:linear_combination");
    // e(A, [tau]G2) == e(B, G2)   (eip4844.c:751)
    TRY(launch_pairing_check(L, d_ok, d_AB + 0, d_AB + 1, nullptr, LINE_G2_TAU, LINE_G2_GEN));
    if (stage_marks) call.mark("stage:pairing");
    call.host_mark("host:t_stage2_enqueued");
    TRY(read_flag(call, d_ok, ok));
    call.host_mark("host:t_verdict_read");
    return RET_OK;
}

int ckzg_b200_verify_kzg_proof(ckzg_b200_ctx* ctx, int* ok, const uint8_t* commitment, const uint8_t* z, const uint8_t* y, const uint8_t* proof) {
    if (!ctx || !ok || !commitment || !z || !y || !proof) return RET_BADARGS;
    *ok = 0;
    Call call(reinterpret_cast<Ctx*>(ctx));
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    uint8_t host[160];
    memcpy(host, commitment, 48);
    memcpy(host + 48, z, 32);
    memcpy(host + 80, y, 32);
    memcpy(host + 112, proof, 48);
    const uint8_t* d_in;
    TRY(call.stage_in(&d_in, host, 160, CKZG_B200_HOST));
    G1Affine *d_cm, *d_pf;
    Fr *d_z, *d_y, *d_r;
    int *d_bad, *d_ok;
    TRY(call.alloc(&d_cm, 1));
    TRY(call.alloc(&d_pf, 1));
    TRY(call.alloc(&d_z, 1));
    TRY(call.alloc(&d_y, 1));
    TRY(call.alloc(&d_r, 1));
    TRY(call.alloc(&d_bad, 1));
    TRY(call.alloc(&d_ok, 1));
    KZG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), call.stream));
    // eip4844.c:317-324: commitment, z, y, proof validated in this order; any failure -> BADARGS
    TRY(launch_g1_validate(L, d_cm, d_in, 1, d_bad, 0));
    TRY(launch_fr_from_bytes(L, d_z, d_in + 48, 32, 1, d_bad));
    TRY(launch_fr_from_bytes(L, d_y, d_in + 80, 32, 1, d_bad));
    TRY(launch_g1_validate(L, d_pf, d_in + 112, 1, d_bad, 0));
    int bad = 0;
    TRY(read_flag(call, d_bad, &bad));
    if (bad) return RET_BADARGS;
    G1* d_AB;
    void* scratch;
    TRY(call.alloc(&d_AB, 2));
    TRY(call.alloc((uint8_t**)&scratch, rlc_scratch_bytes(1)));
    TRY(launch_rlc(L, d_AB, d_cm, d_pf, d_z, d_y, d_r, false, 0, 1, scratch));
    TRY(launch_pairing_check(L, d_ok, d_AB + 0, d_AB + 1, nullptr, LINE_G2_TAU, LINE_G2_GEN));
    TRY(read_flag(call, d_ok, ok));
    return RET_OK;
}

// ---- multi-GPU split -----------------------------------------------------------
// One GLOBAL batch, one Fiat-Shamir challenge, sharded by blob ranges (exact reference semantics for the whole
// batch, eip4844.c:697-765).  A shard keeps its device state -- validated points, the vmsm table its validation left
// behind, z_i, y_i -- between the per-blob stage and the linear combination, so nothing is decompressed or
// validated twice and the combination is the same pair of bucket MSMs as the single-GPU call, with weights
// r^(first + i).  Two front ends share it: the ckzg_b200_verify_shard_* entry points (one process per GPU,
// the exchange steps are the caller's collectives: c-kzg-4844_b200/parallel.py) and the in-library
// multi-device path below (CKZG_B200_DEVICES, exchange through pinned host memory).
}  // extern "C"

namespace kzg {
struct VerifyShard {
    Call call;
    Stage1 s;
    uint64_t n = 0;
    explicit VerifyShard(Ctx* c) : call(c) {}
};

// per-blob stage of a shard; zy_host (n x 64, HOST) receives z || y.  BADARGS for an invalid blob / point.
static int shard_stage1(VerifyShard& sh, uint8_t* zy_host, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs, uint64_t n, int mem) {
    Call& call = sh.call;
    sh.n = n;
    const uint8_t *d_cm, *d_pf;
    TRY(call.stage_in(&d_cm, commitments, n * 48, mem));
    TRY(call.stage_in(&d_pf, proofs, n * 48, mem));
    sh.s.want_shift = true;
    TRY(verify_stage1(call, sh.s, blobs, d_cm, d_pf, n, mem));
    uint8_t* pin = nullptr;
    TRY(call.pin(&pin, 64));
    int* h_bad = (int*)pin;
    KZG_CUDA_TRY(cudaMemcpyAsync(zy_host, sh.s.zy, n * 64, cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaMemcpyAsync(h_bad, sh.s.bad, sizeof(int), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    return *h_bad ? RET_BADARGS : RET_OK;
}
// this shard's share of A and B (2 XYZZ points, 384 bytes of Montgomery limbs: an opaque exchange format) -> HOST
static int shard_stage2(VerifyShard& sh, uint8_t* partial384_host, const uint8_t digest[32], uint64_t first) {
    Call& call = sh.call;
    Launch L = call.launch();
    G1* d_AB;
    uint8_t* scratch;
    TRY(call.alloc(&d_AB, 2));
    TRY(call.alloc(&scratch, rlc_vmsm_scratch_bytes(sh.n)));
    TRY(launch_rlc_vmsm(L, d_AB, sh.s.table, sh.s.z, sh.s.y, digest, first, sh.n, scratch));
    KZG_CUDA_TRY(cudaMemcpyAsync(partial384_host, d_AB, 2 * sizeof(G1), cudaMemcpyDeviceToHost, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));
    return RET_OK;
}
// sum of the shards' partial points, one pairing check
static int shards_finish(Ctx* c, int* ok, const uint8_t* partials, uint64_t n_ranks) {
    Call call(c);
    if (!call.ok) return RET_ERROR;
    Launch L = call.launch();
    static_assert(sizeof(G1) == 192, "exchange format: 2 x 192-byte XYZZ points per shard");
    std::vector<G1> h(2 * n_ranks);
    for (uint64_t k = 0; k < n_ranks; k++) {
        memcpy(&h[k], partials + k * 384, 192);
        memcpy(&h[n_ranks + k], partials + k * 384 + 192, 192);
    }
    G1 *d_a, *d_b, *d_AB;
    int* d_ok;
    TRY(call.alloc(&d_a, 2 * n_ranks + 8));
    TRY(call.alloc(&d_b, 2 * n_ranks + 8));
    TRY(call.alloc(&d_AB, 2));
    TRY(call.alloc(&d_ok, 1));
    KZG_CUDA_TRY(cudaMemcpyAsync(d_a, h.data(), n_ranks * sizeof(G1), cudaMemcpyHostToDevice, call.stream));
    KZG_CUDA_TRY(cudaMemcpyAsync(d_b, h.data() + n_ranks, n_ranks * sizeof(G1), cudaMemcpyHostToDevice, call.stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(call.stream));  // `h` is pageable: the copies above are done with it now
    TRY(launch_g1_sum(L, d_AB + 0, d_a, n_ranks));
    TRY(launch_g1_sum(L, d_AB + 1, d_b, n_ranks));
    TRY(launch_pairing_check(L, d_ok, d_AB + 0, d_AB + 1, nullptr, LINE_G2_TAU, LINE_G2_GEN));
    TRY(read_flag(call, d_ok, ok));
    return RET_OK;
}

namespace {
struct Rendezvous {
    std::mutex m;
    std::condition_variable cv;
    int n, count = 0, gen = 0;
    explicit Rendezvous(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++count == n) {
            count = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};
}  // namespace

// In-library multi-device verify_blob_kzg_proof_batch (HOST pointers): one host thread per device runs the
// per-blob stage of its contiguous range of blobs; z||y meets in pinned host memory, where ONE thread hashes the
// batch transcript (it is hashed on the host on a single GPU too); every device forms its partial sums with the
// weights r^(first + i); device 0 adds the 2 x D partial points and runs the one pairing check.
int verify_blob_batch_multi(Ctx* c, int* ok, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs, uint64_t n, int D) {
    std::vector<uint64_t> first(D + 1);
    for (int d = 0; d <= D; d++) first[d] = n * (uint64_t)d / (uint64_t)D;
    if (getenv("CKZG_B200_DEBUG")) fprintf(stderr, "[ckzg_b200] verify_blob_kzg_proof_batch: %llu blobs over %d devices\n", (unsigned long long)n, D);
    size_t cap = 0;
    uint8_t* pin = (uint8_t*)c->pin_acquire(n * 64 + (size_t)D * 384, &cap);
    if (!pin) return RET_MALLOC;
    uint8_t *zy = pin, *parts = pin + n * 64;
    uint8_t digest[32];
    std::vector<int> rc(D, RET_OK);
    Rendezvous bar(D);
    auto all_ok = [&] {
        for (int d = 0; d < D; d++)
            if (rc[d]) return false;
        return true;
    };
    auto body = [&](int d) {
        const uint64_t f = first[d], cnt = first[d + 1] - first[d];
        VerifyShard sh(c->peers[d]);
        rc[d] = sh.call.ok ? shard_stage1(sh, zy + 64 * f, blobs + f * BLOB_BYTES, commitments + 48 * f, proofs + 48 * f, cnt, CKZG_B200_HOST) : RET_ERROR;
        bar.wait();
        const bool go = all_ok();
        if (go && d == 0) transcript_digest(digest, commitments, zy, proofs, nullptr, n);
        bar.wait();
        if (go) rc[d] = shard_stage2(sh, parts + 384 * (size_t)d, digest, f);
    };
    std::vector<std::thread> th;
    for (int d = 1; d < D; d++) th.emplace_back(body, d);
    body(0);
    for (auto& t : th) t.join();
    int out = RET_OK;
    for (int d = 0; d < D && !out; d++) out = rc[d];  // the reference reports the first invalid input; any BADARGS is BADARGS
    if (!out) out = shards_finish(c, ok, parts, (uint64_t)D);
    c->pin_release(pin, cap);
    return out;
}
}  // namespace kzg

struct ckzg_b200_verify_shard {
    VerifyShard v;
    explicit ckzg_b200_verify_shard(Ctx* c) : v(c) {}
};

extern "C" {

// 160-byte records (C || z || y || proof) of compute_r_powers_for_verify_kzg_proof_batch, eip4844.c:648-660
int ckzg_b200_pack_verify_tuples(uint8_t* tuples, const uint8_t* commitments, const uint8_t* zy, const uint8_t* proofs, uint64_t n) {
    if (!tuples || !commitments || !zy || !proofs) return n ? RET_BADARGS : RET_OK;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t* t = tuples + 160 * i;
        memcpy(t, commitments + 48 * i, 48);
        memcpy(t + 48, zy + 64 * i, 64);
        memcpy(t + 112, proofs + 48 * i, 48);
    }
    return RET_OK;
}

int ckzg_b200_verify_shard_stage1(ckzg_b200_ctx* ctx, ckzg_b200_verify_shard** shard, uint8_t* zy, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs,
                                  uint64_t n_local, int mem) {
    if (!ctx || !shard) return RET_BADARGS;
    *shard = nullptr;
    if (n_local && (!zy || !blobs || !commitments || !proofs)) return RET_BADARGS;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    auto* sh = new (std::nothrow) ckzg_b200_verify_shard(c);
    if (!sh) return RET_MALLOC;
    int rc = sh->v.call.ok ? RET_OK : RET_ERROR;
    if (!rc && n_local) rc = shard_stage1(sh->v, zy, blobs, commitments, proofs, n_local, mem);
    if (rc) {
        delete sh;
        return rc;
    }
    *shard = sh;
    return RET_OK;
}

int ckzg_b200_verify_shard_stage2(ckzg_b200_verify_shard* shard, uint8_t* partial384, const uint8_t* tuples, uint64_t n_total, uint64_t first) {
    if (!shard || !partial384 || (!tuples && n_total) || first + shard->v.n > n_total) return RET_BADARGS;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(shard->v.call.ctx->device) != cudaSuccess) return RET_ERROR;
    int rc = RET_OK;
    if (shard->v.n == 0) {
        memset(partial384, 0, 384);  // zz = zzz = 0: two points at infinity
    } else {
        uint8_t digest[32];
        transcript_digest(digest, nullptr, nullptr, nullptr, tuples, n_total);
        rc = shard_stage2(shard->v, partial384, digest, first);
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

void ckzg_b200_verify_shard_free(ckzg_b200_verify_shard* shard) { delete shard; }

int ckzg_b200_verify_shard_finish(ckzg_b200_ctx* ctx, int* ok, const uint8_t* partials, uint64_t n_ranks) {
    if (!ctx || !ok || !partials || n_ranks == 0) return RET_BADARGS;
    *ok = 0;
    return shards_finish(reinterpret_cast<Ctx*>(ctx), ok, partials, n_ranks);
}

}  // extern "C"
