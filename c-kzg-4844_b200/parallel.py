"""Multi-GPU sharding of verify_blob_kzg_proof_batch: one process per GPU, torch.distributed plumbing.

SURVEY.md section 8(e): the per-blob stage (point validation, Fiat-Shamir challenge z_i, evaluation
y_i) shards freely by blob; the batch challenge r hashes every (C_i, z_i, y_i, proof_i), so the only
data-path exchange is an all-gather of 64 bytes per blob (z_i || y_i) followed by an all-gather of each
rank's two partial linear combinations (2 x 48 bytes) -- the "single small collective" of the
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


def verify_batch_sharded(stage1, stage2, finish, commitments: bytes, proofs: bytes, n_total, device, group=None):
    """One global batch sharded over the process group.

    stage1() -> bytes (n_local x 64: z||y of this rank's blobs; raises on invalid input)
    stage2(tuples: bytes, n_total, first, n_local) -> bytes (144: this rank's partial sums)
    finish(partials: bytes, n_ranks) -> bool
    commitments / proofs: the FULL batch's 48-byte encodings (tiny: replicated on every rank).
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = [shard_range(n_total, r, world) for r in range(world)]
    first, n_local = ranges[rank]
    zy_local = stage1()
    assert len(zy_local) == 64 * n_local
    t = torch.frombuffer(bytearray(zy_local), dtype=torch.uint8).to(device) if n_local else torch.zeros(0, dtype=torch.uint8, device=device)
    zy_all = bytes(_all_gather_bytes(t, [64 * c for _, c in ranges], group).cpu().numpy().tobytes())
    # the 160-byte records the batch challenge hashes (src/eip4844/eip4844.c:648-660)
    tuples = b"".join(
        commitments[48 * i : 48 * i + 48] + zy_all[64 * i : 64 * i + 64] + proofs[48 * i : 48 * i + 48] for i in range(n_total)
    )
    part = stage2(tuples, n_total, first, n_local)
    assert len(part) == 144
    pt = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(device)
    parts = bytes(_all_gather_bytes(pt, [144] * world, group).cpu().numpy().tobytes())
    return finish(parts, world)


def verify_batch_replicas(verify_local, device, group=None):
    """Each rank verifies its own batch; one all-reduce(MIN) combines the verdicts."""
    ok = torch.tensor([1 if verify_local() else 0], dtype=torch.int32, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    return bool(ok.item())
