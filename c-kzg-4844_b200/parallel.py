"""Multi-GPU sharding of verify_blob_kzg_proof_batch: one process per GPU, torch.distributed plumbing.

SURVEY.md section 8(e): the per-blob stage (point validation, Fiat-Shamir challenge z_i, evaluation
y_i) shards freely by blob; the batch challenge r hashes every (C_i, z_i, y_i, proof_i), so the only
data-path exchange is an all-gather of 64 bytes per blob (z_i || y_i) followed by an all-gather of each
rank's two partial linear combinations (2 x 192 bytes, XYZZ) -- the "single small collective" of the
north-star.  NCCL has no elliptic-curve reduction op, hence all-gather + local add instead of
all-reduce.  Any rank (here: every rank, redundantly) finishes with one pairing check.

Two modes:
  * verify_batch_sharded(...)      one global batch of N*n blobs, one challenge, exact reference
                                   semantics for the concatenated batch (two small all-gathers);
  * verify_batch_replicas(...)     each rank verifies its own independent batch and the booleans are
                                   combined with ONE all-reduce(MIN) -- what the reference's own
                                   parallel benchmark does (bindings/go/main_test.go:1037-1101).

The engine entry points are passed in as callables so the host-side logic can be tested on CPU with
the gloo backend and a recording stub (tests/test_parallel_gloo.py).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous block partition: (first, count) of rank `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def _all_gather_bytes(local: torch.Tensor, counts, group=None):
    """All-gather variable-length uint8 tensors (lengths known on every rank) -> one uint8 tensor."""
    world = dist.get_world_size(group)
    width = max(counts)
    pad = torch.zeros(width, dtype=torch.uint8, device=local.device)
    pad[: local.numel()] = local
    bufs = [torch.empty(width, dtype=torch.uint8, device=local.device) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


class ShardError(RuntimeError):
    """Raised on EVERY rank when a stage failed on any rank (code = the largest C_KZG_RET seen)."""

    def __init__(self, code, stage, rank_failed):
        super().__init__("%s failed on rank %d with code %d" % (stage, rank_failed, code))
        self.code, self.stage, self.rank_failed = code, stage, rank_failed


def _agree_on_status(local_exc, stage, device, group=None):
    """All-reduce(MAX) of a status word BEFORE any data collective, so that a stage that raised on one rank
    raises ShardError on all of them instead of leaving the others blocked in the next all_gather.
    The word packs (code, rank) so every rank names the same failing rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    code = 0
    if local_exc is not None:
        code = int(getattr(local_exc, "code", 2) or 2)
    word = torch.tensor([code * 65536 + (rank if code else 0)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(word, op=dist.ReduceOp.MAX, group=group)
    w = int(word.item())
    if w:
        raise ShardError(w // 65536, stage, w % 65536) from local_exc


def verify_batch_sharded(stage1, stage2, finish, commitments: bytes, proofs: bytes, n_total, device, group=None, pack=None):
    """One global batch sharded over the process group.

    stage1() -> bytes (n_local x 64: z||y of this rank's blobs; raises on invalid input)
    stage2(tuples: bytes, n_total, first, n_local) -> bytes (this rank's partial sums; 384 bytes from the engine,
           the same length on every rank)
    finish(partials: bytes, n_ranks) -> bool
    commitments / proofs: the FULL batch's 48-byte encodings (tiny: replicated on every rank).
    pack(commitments, zy, proofs, n) -> bytes: assembles the 160-byte records (the engine's C helper
    ckzg_b200_pack_verify_tuples; default = a pure-Python join, for the CPU tests).

    A stage that raises on any rank (an invalid blob or proof in its shard: C_KZG_BADARGS) makes EVERY rank raise
    ShardError with the same code before the next collective -- no rank is left waiting.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = [shard_range(n_total, r, world) for r in range(world)]
    first, n_local = ranges[rank]
    zy_local, exc = b"", None
    try:
        zy_local = stage1()
        assert len(zy_local) == 64 * n_local
    except Exception as e:  # noqa: BLE001 -- re-raised on every rank below
        exc = e
    _agree_on_status(exc, "stage1", device, group)
    t = torch.frombuffer(bytearray(zy_local), dtype=torch.uint8).to(device) if n_local else torch.zeros(0, dtype=torch.uint8, device=device)
    zy_all = bytes(_all_gather_bytes(t, [64 * c for _, c in ranges], group).cpu().numpy().tobytes())
    # the 160-byte records the batch challenge hashes (src/eip4844/eip4844.c:648-660)
    if pack is not None:
        tuples = pack(commitments, zy_all, proofs, n_total)
    else:
        tuples = b"".join(
            commitments[48 * i : 48 * i + 48] + zy_all[64 * i : 64 * i + 64] + proofs[48 * i : 48 * i + 48] for i in range(n_total)
        )
    part, exc = b"", None
    try:
        part = stage2(tuples, n_total, first, n_local)
    except Exception as e:  # noqa: BLE001
        exc = e
    _agree_on_status(exc, "stage2", device, group)
    pt = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(device)
    parts = bytes(_all_gather_bytes(pt, [len(part)] * world, group).cpu().numpy().tobytes())
    return finish(parts, world)


def verify_batch_replicas(verify_local, device, group=None):
    """Each rank verifies its own batch; one all-reduce(MIN) combines the verdicts."""
    ok = torch.tensor([1 if verify_local() else 0], dtype=torch.int32, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    return bool(ok.item())


# ---- EIP-7594 paths -------------------------------------------------------------------
# compute_cells_and_kzg_proofs / recover_cells_and_kzg_proofs / blob_to_kzg_commitment are independent per blob:
# replicas only, no data-path collective; the helper below just hands every rank its block of blobs and (if asked)
# gathers the fixed-size per-blob outputs.  verify_cell_kzg_proof_batch shards by ranges of (commitment, index,
# cell, proof) tuples: every rank runs an independent batch verification of its range (own Fiat-Shamir challenge)
# and the verdicts meet in ONE all-reduce(MIN) -- the sub-batching the reference's own parallel benchmark uses
# (bindings/go/main_test.go:1037-1101); a batch is valid iff each of its sub-batches is.


def map_blobs_sharded(process_local, n_total, out_bytes_per_blob=0, device="cpu", group=None):
    """process_local(first, count) -> bytes (count x out_bytes_per_blob) for this rank's block of blobs.
    Returns the concatenated outputs of all ranks in blob order when out_bytes_per_blob > 0 (all-gather),
    else this rank's own output."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ranges = [shard_range(n_total, r, world) for r in range(world)]
    first, count = ranges[rank]
    local = process_local(first, count) if count else b""
    if not out_bytes_per_blob or world == 1:
        return local
    assert len(local) == count * out_bytes_per_blob
    t = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(device) if count else torch.zeros(0, dtype=torch.uint8, device=device)
    return bytes(_all_gather_bytes(t, [c * out_bytes_per_blob for _, c in ranges], group).cpu().numpy().tobytes())


def verify_cells_sharded(verify_local, n_tuples, device="cpu", group=None):
    """verify_local(first, count) -> bool for the tuples [first, first + count) (an empty range is valid,
    src/eip7594/eip7594.c:852-855).  One all-reduce(MIN) of the verdicts."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    first, count = shard_range(n_tuples, rank, world)
    ok = torch.tensor([1 if (count == 0 or verify_local(first, count)) else 0], dtype=torch.int32, device=device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    return bool(ok.item())
