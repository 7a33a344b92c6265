/*
 * ckzg_b200.h -- the thin C ABI between host code and the sm_100a CUDA engine.
 *
 * Everything here is `extern "C"`, plain pointers and sizes.  The frozen c-kzg-4844 API
 * (include/ckzg.h: blob_to_kzg_commitment, verify_blob_kzg_proof_batch, ...) is implemented on top of
 * these entry points; they are also what a binding that wants *batched* or *device-resident* calls
 * would bind directly.  Each entry cites the reference interface it replaces
 * (paths relative to the ethereum/c-kzg-4844 tree).
 *
 * Return codes are the reference's C_KZG_RET values (src/common/ret.h:24-29):
 *   0 OK, 1 BADARGS (invalid input bytes), 2 ERROR (CUDA/internal), 3 MALLOC.
 * There is NO CPU fallback: without a usable CUDA device every call returns 2 (C_KZG_ERROR).
 */
#ifndef CKZG_B200_H
#define CKZG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the engine is built with -fvisibility=hidden */
#endif

typedef struct ckzg_b200_ctx ckzg_b200_ctx; /* opaque: device-resident trusted setup + tables */

/* where a caller buffer lives.  HOST buffers need no alignment (the frozen API takes byte pointers from Go slices,
 * Python bytes, ...: they are copied to aligned device staging); DEVICE buffers must be 16-byte aligned (the kernels
 * read field elements as two 128-bit words), which cudaMalloc / torch allocations always are. */
enum { CKZG_B200_HOST = 0, CKZG_B200_DEVICE = 1 };

/*
 * DEVICE-mode ordering contract.  Every engine call runs on its own BLOCKING CUDA stream (plus side streams forked
 * from it), so it is ordered after everything the caller enqueued earlier on the legacy default stream -- the stream
 * torch uses unless told otherwise: `t[...] = x; engine_call(t.data_ptr())` needs no synchronize.  A caller that
 * produces its device buffers on ANOTHER stream (a non-blocking stream, a torch side stream, a per-thread default
 * stream) must either synchronise that stream first or name it here: the calling THREAD's subsequent engine calls
 * then wait for an event recorded on `cuda_stream` (a cudaStream_t) when they start.  NULL restores the default.
 * Calls are synchronous: outputs (HOST or DEVICE) are complete when a call returns, so no ordering is needed after it.
 */
void ckzg_b200_set_caller_stream(void *cuda_stream);

/*
 * Device-side load_trusted_setup (replaces src/setup/setup.c:392-505).
 * Inputs are the three byte arrays of load_trusted_setup(): 4096x48 G1 monomial, 4096x48 G1 Lagrange,
 * 65x96 G2 monomial.  Decompresses on the GPU (no subgroup check, as setup.c:447-477), rejects a
 * monomial-form setup in the Lagrange slot (setup.c:339-358), builds roots of unity, the bit-reversed
 * Lagrange points, the fixed-base window tables for the 4096-point MSM, the FK20 columns and the
 * precomputed pairing lines.  `device` < 0 = current device.
 */
int ckzg_b200_ctx_create(
    ckzg_b200_ctx **out,
    const uint8_t *g1_monomial_bytes,
    const uint8_t *g1_lagrange_bytes,
    const uint8_t *g2_monomial_bytes,
    uint64_t precompute,
    int device
);
void ckzg_b200_ctx_destroy(ckzg_b200_ctx *ctx);
int ckzg_b200_ctx_device(const ckzg_b200_ctx *ctx);
/*
 * One context can span several GPUs of the node: with the environment variable CKZG_B200_DEVICES="0,1,2,3" (or "all")
 * and `device` < 0 -- which is how load_trusted_setup (include/ckzg.h) creates its context -- the returned context owns
 * one replica of the setup per listed device, and every batched entry point below shards HOST-memory batches over them
 * inside the call: contiguous ranges of blobs for commitments / proofs / cells / recovery (no exchange), the per-blob
 * stage + partial sums + one pairing for verify_blob_kzg_proof_batch (one challenge; z||y and the partial points meet in
 * pinned host memory), independently verified sub-batches for verify_cell_kzg_proof_batch.  So a binding that links the
 * frozen API unchanged drives all listed GPUs.  DEVICE-memory inputs always run on the device that owns them.
 */
int ckzg_b200_ctx_device_count(const ckzg_b200_ctx *ctx);
/*
 * Device memory.  A context holds ~40 MB of setup data.  Two fixed-base tables are built on FIRST USE of the APIs that
 * need them and sized from the HBM that is free on the device at that moment (verification-only users build neither):
 *   commitments / proofs (blob_to_kzg_commitment, compute_*_proof): window 14 -> 61 GB | 13 -> 32 GB | 12 -> 18 GB |
 *       none (a 9 MB table + bucket MSM) when less than 60 GB are free;   CKZG_B200_COMMIT_WINDOW = 0 | 10..14 pins it;
 *   FK20 cell proofs (compute_cells_and_kzg_proofs, recover_*):          window 12 -> 35 GB | 10 -> 10.5 GB | 8 -> 3.2 GB;
 *                                                                        CKZG_B200_FK_WINDOW = 8 | 10 | 12 pins it.
 * Results are identical for every choice.  out = { commitment table bytes (0 = not built / bucket form), its window,
 * FK20 table bytes (0 = not built), its window, the window a commitment-table build would choose NOW, same for FK20 }.
 */
int ckzg_b200_ctx_table_info(const ckzg_b200_ctx *ctx, uint64_t out[6]);

/*
 * Batched blob_to_kzg_commitment (src/eip4844/eip4844.c:264 applied to n blobs).
 *   blobs:  n x 131072 bytes, out: n x 48 bytes, both in `mem` space.
 *   status (optional, HOST memory, n ints): per-blob C_KZG_RET (BADARGS for a non-canonical element).
 * Returns OK if every blob was valid, else the first non-OK status.
 */
int ckzg_b200_blob_to_kzg_commitment_batch(
    ckzg_b200_ctx *ctx, uint8_t *out, const uint8_t *blobs, uint64_t n, int mem, int *status
);

/*
 * Batched compute_blob_kzg_proof (src/eip4844/eip4844.c:506) and compute_kzg_proof (:382).
 *   commitments: n x 48.  proofs out: n x 48.
 *   compute_kzg_proof: zs n x 32 in, ys n x 32 out.
 */
int ckzg_b200_compute_blob_kzg_proof_batch(
    ckzg_b200_ctx *ctx, uint8_t *proofs, const uint8_t *blobs, const uint8_t *commitments, uint64_t n, int mem, int *status
);
int ckzg_b200_compute_kzg_proof_batch(
    ckzg_b200_ctx *ctx, uint8_t *proofs, uint8_t *ys, const uint8_t *blobs, const uint8_t *zs, uint64_t n, int mem, int *status
);

/*
 * verify_blob_kzg_proof_batch (src/eip4844/eip4844.c:775) with the reference's exact semantics:
 * n == 0 -> ok, n == 1 -> the single-blob equation, else one random linear combination with the
 * Fiat-Shamir challenge of :597-680 and one pairing check.
 */
int ckzg_b200_verify_blob_kzg_proof_batch(
    ckzg_b200_ctx *ctx, int *ok, const uint8_t *blobs, const uint8_t *commitments, const uint8_t *proofs, uint64_t n, int mem
);

/* verify_kzg_proof (src/eip4844/eip4844.c:302): one (commitment, z, y, proof) tuple, host pointers. */
int ckzg_b200_verify_kzg_proof(
    ckzg_b200_ctx *ctx, int *ok, const uint8_t *commitment, const uint8_t *z, const uint8_t *y, const uint8_t *proof
);

/*
 * Multi-GPU split of verify_blob_kzg_proof_batch for one process per GPU: ONE global batch of
 * n_total blobs, one Fiat-Shamir challenge, exact reference semantics for the whole batch (eip4844.c:697-765).
 *   stage1: the per-blob stage of this rank's n_local blobs (validation, z_i, y_i).  zy out = n_local x 64 bytes
 *           (z || y, canonical big-endian), HOST memory.  The shard object keeps the validated points, the table
 *           their validation left behind, z and y ON THE DEVICE for stage 2.  BADARGS (no shard) for an invalid input.
 *   (caller: all-gather zy, 64 B per blob; ckzg_b200_pack_verify_tuples builds the 160-byte records)
 *   stage2: tuples = n_total x 160 bytes (C48 || z32 || y32 || proof48), HOST memory, hashed here into the challenge r
 *           (eip4844.c:612-668); this rank owns [first, first + n_local) and forms its share of
 *           A = sum r^i proof_i and B = sum r^i z_i proof_i + sum r^i C_i - [sum r^i y_i]G as one pair of bucket
 *           MSMs with weights r^(first + i).  partial out = 384 bytes (two XYZZ points; an opaque exchange format).
 *   (caller: all-gather the partials, 384 B per rank -- NCCL has no elliptic-curve reduction op)
 *   finish: partials = n_ranks x 384 bytes, HOST memory: sums them, one pairing check.
 * A shard must be freed (verify_shard_free) by the thread's process that created it; n_local may be 0.
 * The same code runs inside the library when a context spans several devices (ckzg_b200_ctx_create, CKZG_B200_DEVICES).
 */
typedef struct ckzg_b200_verify_shard ckzg_b200_verify_shard;
int ckzg_b200_verify_shard_stage1(
    ckzg_b200_ctx *ctx, ckzg_b200_verify_shard **shard, uint8_t *zy, const uint8_t *blobs, const uint8_t *commitments,
    const uint8_t *proofs, uint64_t n_local, int mem
);
/* host helper between the stages: tuples[i] = commitments[i] || zy[i] || proofs[i] (160 B).  All HOST memory. */
int ckzg_b200_pack_verify_tuples(uint8_t *tuples, const uint8_t *commitments, const uint8_t *zy, const uint8_t *proofs, uint64_t n);
int ckzg_b200_verify_shard_stage2(ckzg_b200_verify_shard *shard, uint8_t *partial384, const uint8_t *tuples, uint64_t n_total, uint64_t first);
void ckzg_b200_verify_shard_free(ckzg_b200_verify_shard *shard);
int ckzg_b200_verify_shard_finish(ckzg_b200_ctx *ctx, int *ok, const uint8_t *partials, uint64_t n_ranks);

/*
 * EIP-7594.  Batched compute_cells_and_kzg_proofs (src/eip7594/eip7594.c:61): per blob 128 cells
 * (2048 B each, bit-reversed evaluation order) and/or 128 FK20 cell proofs (src/eip7594/fk20.c:139).
 * Either output may be NULL, not both (eip7594.c:72-74).
 */
int ckzg_b200_compute_cells_and_kzg_proofs_batch(
    ckzg_b200_ctx *ctx, uint8_t *cells, uint8_t *proofs, const uint8_t *blobs, uint64_t n, int mem, int *status
);
/*
 * Batched recover_cells_and_kzg_proofs (src/eip7594/eip7594.c:177): every blob of the batch comes
 * with the same number `num_cells` of (index, cell) pairs; indices are n x num_cells uint64 in HOST
 * memory (validated on the host exactly as eip7594.c:191-213), cells n x num_cells x 2048 B in `mem`.
 * recovered_proofs may be NULL.
 */
int ckzg_b200_recover_cells_and_kzg_proofs_batch(
    ckzg_b200_ctx *ctx, uint8_t *recovered_cells, uint8_t *recovered_proofs, const uint64_t *cell_indices,
    const uint8_t *cells, uint64_t num_cells, uint64_t n, int mem, int *status
);
/* verify_cell_kzg_proof_batch (src/eip7594/eip7594.c:825); cell_indices in HOST memory. */
int ckzg_b200_verify_cell_kzg_proof_batch(
    ckzg_b200_ctx *ctx, int *ok, const uint8_t *commitments, const uint64_t *cell_indices, const uint8_t *cells,
    const uint8_t *proofs, uint64_t num_cells, int mem
);

/*
 * Per-blob entry points with call coalescing: what the frozen per-blob API
 * (blob_to_kzg_commitment src/eip4844/eip4844.c:264, compute_kzg_proof :382, compute_blob_kzg_proof :506,
 * compute_cells_and_kzg_proofs src/eip7594/eip7594.c:61, recover_cells_and_kzg_proofs :177) calls.  Host
 * pointers, synchronous, re-entrant: callers that arrive while a batch is running are merged into the next
 * batched engine call; a lone caller runs at once.  Results and return codes per caller are those of the
 * corresponding n = 1 call.
 */
int ckzg_b200_blob_to_kzg_commitment_coalesced(ckzg_b200_ctx *ctx, uint8_t *out48, const uint8_t *blob);
int ckzg_b200_compute_blob_kzg_proof_coalesced(ckzg_b200_ctx *ctx, uint8_t *proof48, const uint8_t *blob, const uint8_t *commitment48);
int ckzg_b200_compute_kzg_proof_coalesced(ckzg_b200_ctx *ctx, uint8_t *proof48, uint8_t *y32, const uint8_t *blob, const uint8_t *z32);
int ckzg_b200_compute_cells_and_kzg_proofs_coalesced(ckzg_b200_ctx *ctx, uint8_t *cells, uint8_t *proofs, const uint8_t *blob);
int ckzg_b200_recover_cells_and_kzg_proofs_coalesced(
    ckzg_b200_ctx *ctx, uint8_t *recovered_cells, uint8_t *recovered_proofs, const uint64_t *cell_indices, const uint8_t *cells, uint64_t num_cells
);
/* {requests, batches, largest batch} x {commitment, blob proof, kzg proof, cells, recover} since context creation */
int ckzg_b200_coalesce_stats(ckzg_b200_ctx *ctx, uint64_t out15[15]);
/* on = 0: every call runs alone; default on (environment CKZG_B200_COALESCE=0 disables at context creation) */
int ckzg_b200_coalesce_enable(ckzg_b200_ctx *ctx, int on);

/* Measurement hook: `threads` native host threads call per-blob entry point `op` (0 = blob_to_kzg_commitment,
 * 1 = compute_cells_and_kzg_proofs) `reps` times each on blobs from a HOST array; *seconds = wall time. */
int ckzg_b200_bench_per_blob_callers(ckzg_b200_ctx *ctx, int op, int threads, int reps, const uint8_t *blobs, uint64_t n_blobs, double *seconds);

/* Internal Fiat-Shamir challenge, exposed for the vectors in tests/compute_challenge
 * (src/eip4844/eip4844.c:147; commitment given in its canonical 48-byte form). out = 32 bytes BE.
 * ctx may be NULL (the hash needs no setup; the current CUDA device is used). */
int ckzg_b200_compute_challenge(ckzg_b200_ctx *ctx, uint8_t *out32, const uint8_t *blob, const uint8_t *commitment48);

/* hash_to_bls_field (src/common/bytes.c:123): 32-byte digest -> canonical field element bytes (mod r on the device) */
int ckzg_b200_hash_to_bls_field(uint8_t *out32, const uint8_t *digest32);
/* validate_kzg_g1 (src/common/bytes.c:81) for one 48-byte point, no context needed: *ok = 1 if valid */
int ckzg_b200_validate_g1(int *ok, const uint8_t *p48);

/* Counters for bench.py: kernels launched by this library since the context was created. */
uint64_t ckzg_b200_launch_count(const ckzg_b200_ctx *ctx);

/* Device timing for bench.py.  level 0 = off; 1 = whole-call begin/end events only (concurrent stages
 * stay concurrent); 2 = per-kernel events, stages serialised on one stream so the attribution is exact.
 * Enabling also resets the counters.  Run calls, then dump a
 * JSON object {"calls": n, "call_ms": total, "kernels": {name: [total_ms, launches], ...}} into buf.
 * Times come from CUDA events recorded on the stream each call launches on. Returns bytes written. */
void ckzg_b200_profile_enable(ckzg_b200_ctx *ctx, int level);
int ckzg_b200_profile_dump(ckzg_b200_ctx *ctx, char *buf, size_t cap);

/* Device self-tests used by tests/test_gpu_units.py: run `op` over n operand pairs.
 *   op 0: Fp mul, 1: Fp add, 2: Fp sub, 3: Fp inv(a), 4: Fr mul, 5: Fr inv(a), 6: Fp sqr
 *   operands/results are plain little-endian 32-bit limbs (12 per Fp, 8 per Fr), HOST memory. */
int ckzg_b200_selftest_field(int op, uint32_t *out, const uint32_t *a, const uint32_t *b, uint64_t n);
/* Fp Montgomery-multiplier throughput probe (the measured integer-pipe roofline denominator):
 * blocks x threads threads each run `iters` rounds of `ilp` (1, 2 or 4) independent dependent-chains.
 * ilp bits 8..15: active lanes per warp (0 = 32); bit 16: FP64 FMA probe (8 FMAs per thread per round). */
int ckzg_b200_selftest_mulbench(int ilp, int iters, int blocks, int threads, float *ms_out);
/* Arms (dev_buf != NULL) or disarms the placement probe of the stage-1 kernels: dev_buf is a DEVICE array
 * of uint32, [0] = record count (set to 0), [1] = capacity, records follow:
 * kernel id << 28 | warp in block << 24 | %smid << 8 | %warpid.  Test/measurement hook. */
int ckzg_b200_debug_placement(uint32_t *dev_buf);
/* Arms (or, with NULL, disarms) the device-side timeline probe of the stage-1 kernels: dev_buf[0] = record count (set to 0),
 * [1] = capacity, then (kernel id, low 32 bits of %globaltimer in ns) pairs written by CTA 0 at kernel start and end
 * (end: id | 0x80; hash kernels carry their first SHA block in bits 8+).  Measurement hook (tools/e2e_probe.py). */
int ckzg_b200_debug_timers(uint32_t *dev_buf);
/* Measurement hook (tools/pairing_probe.py): one pairing check e(-P0, G2[0]) e(P1, G2[1]) on the two compressed G1 points
 * given (HOST memory) with clock64() marks -- ticks64[0..9] = start, inputs ready, Miller loops done, product of the two
 * Miller values, before / after the Fp12 inversion, easy part done, first exponentiation by the curve parameter done, hard
 * part done, verdict -- followed by `reps` repetitions of each cooperative tower operation: ticks64[16 + 2k], [17 + 2k] =
 * start / end for k = cyclotomic square, product, square, line product, conjugation, copy, Frobenius, inversion (2 runs). */
int ckzg_b200_debug_pairing_probe(ckzg_b200_ctx *ctx, long long *ticks64, int *ok, const uint8_t *two_g1_48, int reps);
/* Measurement hook: best-of-`reps` wall time (ms) of uploading `bytes` from HOST memory the way every HOST-mode call does
 * (pageable sources: pinned staging ring filled by the context's host threads, CKZG_B200_HOST_THREADS /
 * CKZG_B200_STAGE_SLOT_MB; pinned sources: one DMA).  mode 1: cudaHostRegister + direct DMA + cudaHostUnregister. */
int ckzg_b200_debug_upload(ckzg_b200_ctx *ctx, const uint8_t *host, uint64_t bytes, int reps, int mode, double *ms_best);
/* op 0: [k]P (+Q), op 1: validate (subgroup), op 2: uncompress only; compressed points, HOST memory.
 * ok_out[i] = 1 if the input decoded/validated.  k = 8 limbs per scalar. */
int ckzg_b200_selftest_g1(int op, uint8_t *out48, int *ok_out, const uint8_t *p48, const uint32_t *k, const uint8_t *q48, uint64_t n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CKZG_B200_H */
