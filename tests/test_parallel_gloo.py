"""CPU (gloo, world_size 2 and 3) test of the multi-GPU host logic in c-kzg-4844_b200/parallel.py with a
recording stub in place of the engine: shard ranges tile the batch, every rank assembles the SAME
160-byte transcript in blob order, partial sums are all-gathered in rank order, verdicts all-reduce."""
import hashlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as entry


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _zy(i):
    return hashlib.sha256(b"zy%d" % i).digest() * 2


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    entry.load_package()
    import importlib

    par = importlib.import_module("ckzg_b200.parallel")
    cms = b"".join(hashlib.sha256(b"c%d" % i).digest()[:24] * 2 for i in range(n_total))
    prs = b"".join(hashlib.sha256(b"p%d" % i).digest()[:24] * 2 for i in range(n_total))
    first, cnt = par.shard_range(n_total, rank, world)
    seen = {}

    def stage1():
        return b"".join(_zy(first + k) for k in range(cnt))

    def stage2(tuples, n, f, c):
        seen["tuples"] = hashlib.sha256(tuples).hexdigest()
        assert (n, f, c) == (n_total, first, cnt)
        return bytes([rank]) * 144

    def finish(parts, nr):
        seen["parts"] = parts
        return nr == world and all(parts[144 * r : 144 * r + 144] == bytes([r]) * 144 for r in range(world))

    ok = par.verify_batch_sharded(stage1, stage2, finish, cms, prs, n_total, torch.device("cpu"))
    want = hashlib.sha256(b"".join(cms[48 * i : 48 * i + 48] + _zy(i) + prs[48 * i : 48 * i + 48] for i in range(n_total))).hexdigest()
    rep_all_true = par.verify_batch_replicas(lambda: True, torch.device("cpu"))
    rep_one_false = par.verify_batch_replicas(lambda: rank != world - 1, torch.device("cpu"))
    q.put((rank, ok, seen["tuples"] == want, first, cnt, rep_all_true, rep_one_false))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 7), (3, 8), (2, 1)])
def test_sharded_verify_host_logic(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    covered = []
    for rank, ok, same, first, cnt, rep_true, rep_false in res:
        assert ok and same and rep_true and not rep_false
        covered += list(range(first, first + cnt))
    assert covered == list(range(n_total))


def _worker_failing(rank, world, port, n_total, fail_stage, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import datetime

    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=60))
    entry.load_package()
    import importlib

    par = importlib.import_module("ckzg_b200.parallel")
    first, cnt = par.shard_range(n_total, rank, world)

    def bad_args():
        e = ValueError("C_KZG_BADARGS")
        e.code = 1
        raise e

    def stage1():
        if fail_stage == 1 and rank == world - 1:
            bad_args()
        return b"\0" * (64 * cnt)

    def stage2(tuples, n, f, c):
        if fail_stage == 2 and rank == 0:
            bad_args()
        return b"\0" * 144

    try:
        par.verify_batch_sharded(stage1, stage2, lambda parts, nr: True, b"\0" * (48 * n_total), b"\0" * (48 * n_total), n_total, torch.device("cpu"))
        q.put((rank, "returned", None, None))
    except par.ShardError as e:
        q.put((rank, "ShardError", e.code, e.rank_failed))
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_stage", [1, 2])
def test_sharded_verify_failure_on_one_rank_raises_everywhere(fail_stage):
    """a stage raising on one rank used to leave the others blocked in all_gather.  Now every rank
    raises ShardError with the failing rank's C_KZG_RET before the next collective."""
    world, n_total = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_failing, args=(r, world, port, n_total, fail_stage, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    failing = world - 1 if fail_stage == 1 else 0
    assert res == [(r, "ShardError", 1, failing) for r in range(world)]


def test_shard_range_properties():
    entry.load_package()
    import importlib

    par = importlib.import_module("ckzg_b200.parallel")
    for n in (0, 1, 5, 64, 4096, 4097):
        for w in (1, 2, 3, 4, 8):
            rs = [par.shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in rs) == n
            assert all(rs[i][0] + rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in rs) - min(c for _, c in rs) <= 1


def _cells_worker(rank, world, port, q):
    """EIP-7594 sharding with the unmodified reference (oracle/_ref) as every rank's engine: the sharded verdict must
    equal the reference's verdict on the whole batch, for a valid batch and for one with a single bad proof."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    entry.load_package()
    import importlib

    from oracle import ref_lib
    from gpu_common import synth_blob

    par = importlib.import_module("ckzg_b200.parallel")
    ref = ref_lib.CKZG()
    nblobs = 3
    blobs = [synth_blob(700 + b) for b in range(nblobs)]

    # replicas: every rank computes the cells + proofs of its block of blobs; outputs gathered in blob order
    def process(first, count):
        out = b""
        for b in range(first, first + count):
            cells, proofs = ref.compute_cells_and_kzg_proofs(blobs[b])
            out += cells + proofs
        return out

    per_blob = 128 * 2048 + 128 * 48
    gathered = par.map_blobs_sharded(process, nblobs, per_blob)
    assert len(gathered) == nblobs * per_blob
    cms = [ref.blob_to_kzg_commitment(b) for b in blobs]
    tuples = []  # row-major by blob, every third cell
    for b in range(nblobs):
        blk = gathered[b * per_blob : (b + 1) * per_blob]
        for k in range(0, 128, 3):
            tuples.append((cms[b], k, blk[2048 * k : 2048 * (k + 1)], blk[128 * 2048 + 48 * k : 128 * 2048 + 48 * (k + 1)]))

    def verdict(ts):
        def local(first, count):
            part = ts[first : first + count]
            return ref.verify_cell_kzg_proof_batch(b"".join(t[0] for t in part), [t[1] for t in part], b"".join(t[2] for t in part), b"".join(t[3] for t in part))

        return par.verify_cells_sharded(local, len(ts))

    good = verdict(tuples)
    bad_tuples = list(tuples)
    j = len(tuples) - 2  # lands in the last rank's range
    bad_tuples[j] = (bad_tuples[j][0], bad_tuples[j][1], bad_tuples[j][2], tuples[j - 1][3])
    bad = verdict(bad_tuples)
    whole_good = ref.verify_cell_kzg_proof_batch(b"".join(t[0] for t in tuples), [t[1] for t in tuples], b"".join(t[2] for t in tuples), b"".join(t[3] for t in tuples))
    whole_bad = ref.verify_cell_kzg_proof_batch(b"".join(t[0] for t in bad_tuples), [t[1] for t in bad_tuples], b"".join(t[2] for t in bad_tuples), b"".join(t[3] for t in bad_tuples))
    q.put((rank, good, bad, whole_good, whole_bad, hashlib.sha256(gathered).hexdigest()))
    dist.destroy_process_group()


def test_cells_sharding_matches_reference_verdicts():
    from oracle import ref_lib

    if not os.path.exists(ref_lib.REF_SO):
        pytest.skip("oracle/_ref not built")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cells_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert len({r[5] for r in res}) == 1  # every rank holds the same gathered outputs
    for rank, good, bad, whole_good, whole_bad, _ in res:
        assert good is True and whole_good is True
        assert bad is False and whole_bad is False
