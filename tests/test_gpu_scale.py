"""-m gpu: BASELINE.json sizes through size-independent properties and sampled differential checks
(the consensus vectors stop at 7 blobs / 128 cells)."""
import os
import random
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def env():
    import torch

    import __graft_entry__ as entry
    import bench

    mod = entry.load_package()
    ts = mod.load_trusted_setup()
    n = 1024
    host = torch.from_numpy(bench.synth_blobs(n, 31337))
    dev = host.cuda()
    cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
    mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
    return mod, ts, n, host, dev, cms, prs


def test_batch1024_commitments_and_proofs_sampled_vs_reference(env):
    from oracle import ref_lib

    mod, ts, n, host, dev, cms, prs = env
    if not os.path.exists(ref_lib.REF_SO):
        pytest.skip("oracle/_ref not built")
    ref = ref_lib.CKZG()
    hb, hc, hp = host.numpy().tobytes(), cms.cpu().numpy().tobytes(), prs.cpu().numpy().tobytes()
    for i in random.Random(5).sample(range(n), 12) + [0, n - 1]:
        blob = hb[131072 * i : 131072 * (i + 1)]
        c = ref.blob_to_kzg_commitment(blob)
        assert hc[48 * i : 48 * i + 48] == c, i
        assert hp[48 * i : 48 * i + 48] == ref.compute_blob_kzg_proof(blob, c), i


def test_batch1024_verify_true_and_single_corruption_false(env):
    mod, ts, n, host, dev, cms, prs = env
    assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts) is True
    # host-pointer call (chunked upload path) must agree
    assert mod.verify_blob_kzg_proof_batch(host.numpy().tobytes(), cms.cpu().numpy().tobytes(), prs.cpu().numpy().tobytes(), ts) is True
    bad = prs.clone()
    i = 777
    bad[48 * i : 48 * i + 48] = prs[48 * (i + 1) : 48 * (i + 2)].clone()
    assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), bad.data_ptr(), n, ts) is False
    # one flipped blob byte (still canonical): the evaluation changes -> false
    blobs2 = dev.clone()
    blobs2[131072 * 500 + 31] ^= 1
    assert mod.verify_blob_kzg_proof_batch_device(blobs2.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts) is False


def test_cells_recover_roundtrip_batch(env):
    """encode -> erase half -> decode at batch 128 (BASELINE configs[2]/[3] shapes), plus cell-proof verification."""
    import torch

    mod, ts, n, host, dev, cms, prs = env
    m = 128
    cells = torch.empty(m * 262144, dtype=torch.uint8, device="cuda")
    cprf = torch.empty(m * 128 * 48, dtype=torch.uint8, device="cuda")
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), cprf.data_ptr(), dev.data_ptr(), m, ts)
    # first half of the extension is the blob itself
    assert torch.equal(cells.view(m, 2, 131072)[:, 0, :].reshape(-1), dev[: m * 131072])
    for pattern in (list(range(0, 128, 2)), list(range(64, 128)), sorted(random.Random(9).sample(range(128), 70))):
        idx = pattern * m
        given = cells.view(m, 128, 2048)[:, pattern, :].contiguous()
        rc = torch.empty_like(cells)
        rp = torch.empty_like(cprf)
        mod.recover_cells_and_kzg_proofs_device(rc.data_ptr(), rp.data_ptr(), idx, given.data_ptr(), len(pattern), m, ts)
        assert torch.equal(rc, cells) and torch.equal(rp, cprf), pattern[:3]
    # cell proofs of 4 blobs verify against the blob commitments; a swapped cell does not
    hc, hp, hcm = cells.cpu().numpy().tobytes(), cprf.cpu().numpy().tobytes(), cms.cpu().numpy().tobytes()
    sel = [(b, k) for b in range(4) for k in range(0, 128, 3)]
    cm_l = [hcm[48 * b : 48 * b + 48] for b, k in sel]
    idx_l = [k for b, k in sel]
    cell_l = [hc[(b * 128 + k) * 2048 : (b * 128 + k + 1) * 2048] for b, k in sel]
    prf_l = [hp[(b * 128 + k) * 48 : (b * 128 + k + 1) * 48] for b, k in sel]
    assert mod.verify_cell_kzg_proof_batch(cm_l, idx_l, cell_l, prf_l, ts) is True
    cell_l[5], cell_l[6] = cell_l[6], cell_l[5]
    assert mod.verify_cell_kzg_proof_batch(cm_l, idx_l, cell_l, prf_l, ts) is False


def test_verify_cells_batch8192_and_controls(env):
    """verify_cell_kzg_proof_batch at 64 blobs x 128 cells (row-major by blob, bindings/go/main_test.go:1016-1028):
    the bucket-MSM form must accept the engine's own (parity-checked) cells + proofs, in any order and with
    duplicates, reject one wrong proof / cell / index, and report invalid encodings as BADARGS."""
    import torch

    from oracle import ref_lib

    mod, ts, n, host, dev, cms, prs = env
    m = 64
    cells = torch.empty(m * 262144, dtype=torch.uint8, device="cuda")
    cprf = torch.empty(m * 128 * 48, dtype=torch.uint8, device="cuda")
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), cprf.data_ptr(), dev.data_ptr(), m, ts)
    hc, hp, hcm = cells.cpu().numpy().tobytes(), cprf.cpu().numpy().tobytes(), cms.cpu().numpy().tobytes()
    cm_l = [hcm[48 * b : 48 * b + 48] for b in range(m) for k in range(128)]
    idx_l = [k for b in range(m) for k in range(128)]
    cell_l = [hc[i * 2048 : (i + 1) * 2048] for i in range(m * 128)]
    prf_l = [hp[i * 48 : (i + 1) * 48] for i in range(m * 128)]
    assert mod.verify_cell_kzg_proof_batch(cm_l, idx_l, cell_l, prf_l, ts) is True
    # shuffled, with duplicates (eip7594.c:345-376: dedup order follows first appearance)
    order = list(range(m * 128))
    random.Random(4).shuffle(order)
    order = order[:3000] + order[:17]
    pick = lambda l: [l[i] for i in order]
    assert mod.verify_cell_kzg_proof_batch(pick(cm_l), pick(idx_l), pick(cell_l), pick(prf_l), ts) is True
    if os.path.exists(ref_lib.REF_SO):  # the reference agrees on a 300-cell sample of the same data
        ref = ref_lib.CKZG()
        sub = order[:300]
        assert ref.verify_cell_kzg_proof_batch(b"".join(cm_l[i] for i in sub), [idx_l[i] for i in sub], b"".join(cell_l[i] for i in sub), b"".join(prf_l[i] for i in sub)) is True
    # negative controls
    p2 = list(prf_l)
    p2[4097], p2[4098] = p2[4098], p2[4097]
    assert mod.verify_cell_kzg_proof_batch(cm_l, idx_l, cell_l, p2, ts) is False
    c2 = list(cell_l)
    c2[8000] = c2[8000][:2047] + bytes([c2[8000][2047] ^ 1])
    assert mod.verify_cell_kzg_proof_batch(cm_l, idx_l, c2, prf_l, ts) is False
    i2 = list(idx_l)
    i2[77] = (i2[77] + 64) % 128
    assert mod.verify_cell_kzg_proof_batch(cm_l, i2, cell_l, prf_l, ts) is False
    k2 = list(cm_l)
    k2[128 * 9 + 3] = cm_l[0]
    assert mod.verify_cell_kzg_proof_batch(k2, idx_l, cell_l, prf_l, ts) is False
    # invalid encodings
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    c3 = list(cell_l)
    c3[5000] = c3[5000][:64] + R.to_bytes(32, "big") + c3[5000][96:]
    with pytest.raises(Exception):
        mod.verify_cell_kzg_proof_batch(cm_l, idx_l, c3, prf_l, ts)
    p3 = list(prf_l)
    p3[8191] = bytes(48)
    with pytest.raises(Exception):
        mod.verify_cell_kzg_proof_batch(cm_l, idx_l, cell_l, p3, ts)


def test_concurrent_callers_share_one_settings(env):
    """The reference allows many threads on one const KZGSettings (bindings/rust/src/bindings/mod.rs:912,
    bindings/go/main_test.go:957-970): every call here owns its stream and pool allocations."""
    import threading

    mod, ts, n, host, dev, cms, prs = env
    hb, hc = host.numpy().tobytes(), cms.cpu().numpy().tobytes()
    hp = prs.cpu().numpy().tobytes()
    errors = []

    def worker(w):
        try:
            for k in range(6):
                i = (w * 6 + k) % 64
                blob = hb[131072 * i : 131072 * (i + 1)]
                assert mod.blob_to_kzg_commitment(blob, ts) == hc[48 * i : 48 * i + 48]
                assert mod.verify_blob_kzg_proof(blob, hc[48 * i : 48 * i + 48], hp[48 * i : 48 * i + 48], ts) is True
                if k % 3 == 0:
                    assert mod.verify_blob_kzg_proof_batch(hb[: 131072 * 8], hc[: 48 * 8], hp[: 48 * 8], ts) is True
        except Exception as e:  # noqa: BLE001
            errors.append((w, repr(e)))

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_coalesced_per_blob_api_matches_batched_results(env):
    """concurrent callers of the frozen per-blob API are merged into batched engine calls.  Every
    caller must get exactly the bytes (and the return code) its own n = 1 call would have produced."""
    import threading

    import torch

    mod, ts, n, host, dev, cms, prs = env
    m = 24
    hb, hc, hp = host.numpy().tobytes(), cms.cpu().numpy().tobytes(), prs.cpu().numpy().tobytes()
    cells = torch.empty(m * 262144, dtype=torch.uint8, device="cuda")
    cprf = torch.empty(m * 128 * 48, dtype=torch.uint8, device="cuda")
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), cprf.data_ptr(), dev.data_ptr(), m, ts)
    want_cells, want_cprf = cells.cpu().numpy().tobytes(), cprf.cpu().numpy().tobytes()
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    bad_blob = R.to_bytes(32, "big") + hb[32:131072]
    before = mod.coalesce_stats(ts)
    errors = []

    def worker(i):
        try:
            blob = hb[131072 * i : 131072 * (i + 1)]
            cm, pf = hc[48 * i : 48 * i + 48], hp[48 * i : 48 * i + 48]
            for _ in range(2):
                assert mod.blob_to_kzg_commitment(blob, ts) == cm
                assert mod.compute_blob_kzg_proof(blob, cm, ts) == pf
                c, p = mod.compute_cells_and_kzg_proofs(blob, ts)
                assert b"".join(c) == want_cells[262144 * i : 262144 * (i + 1)]
                assert b"".join(p) == want_cprf[6144 * i : 6144 * (i + 1)]
                idx = list(range(i % 2, 128, 2)) if i % 3 else list(range(60, 128))
                rc, rp = mod.recover_cells_and_kzg_proofs(idx, [c[k] for k in idx], ts)
                assert rc == c and rp == p
                if i % 5 == 0:  # an invalid blob inside a batch fails alone
                    with pytest.raises(Exception):
                        mod.blob_to_kzg_commitment(bad_blob, ts)
                    with pytest.raises(Exception):
                        mod.compute_cells_and_kzg_proofs(bad_blob, ts)
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(m)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    after = mod.coalesce_stats(ts)
    merged = {k: (after[k][0] - before[k][0], after[k][1] - before[k][1], after[k][2]) for k in after}
    # 24 concurrent callers: far fewer engine calls than requests for the long-running operations
    assert merged["compute_cells_and_kzg_proofs"][0] >= 2 * m and merged["compute_cells_and_kzg_proofs"][1] < merged["compute_cells_and_kzg_proofs"][0]
    assert merged["compute_cells_and_kzg_proofs"][2] > 1
    # switched off, the same calls still work (every call alone)
    mod.coalesce_enable(ts, False)
    try:
        assert mod.blob_to_kzg_commitment(hb[:131072], ts) == hc[:48]
    finally:
        mod.coalesce_enable(ts, True)


def test_fk20_affine_batch_matches_single_blob_path(env):
    """Batches of >= 8 blobs add the FK20 table points pairwise in affine coordinates with batched inversions
    (msm_affine.cu); single blobs use XYZZ accumulators (fk20.cu).  Both must give the same bytes, also for blobs
    whose digits are mostly zero (infinity operands), all equal, or extreme."""
    import torch

    mod, ts, n, host, dev, cms, prs = env
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    z, one, top = (0).to_bytes(32, "big"), (1).to_bytes(32, "big"), (R - 1).to_bytes(32, "big")
    hb = host.numpy().tobytes()
    special = [
        z * 4096,
        one * 4096,
        top * 4096,
        z * 4000 + top + z * 95,
        b"".join(i.to_bytes(32, "big") for i in range(4096)),
        (one + top) * 2048,
        z * 64 + one * 64 + z * 3968,
        b"".join(((1 << 254) + 7 * i).to_bytes(32, "big") for i in range(4096)),
    ]
    blobs = special + [hb[131072 * i : 131072 * (i + 1)] for i in range(8)]
    m = len(blobs)
    d_in = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
    cells = torch.empty(m * 262144, dtype=torch.uint8, device="cuda")
    cprf = torch.empty(m * 128 * 48, dtype=torch.uint8, device="cuda")
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), cprf.data_ptr(), d_in.data_ptr(), m, ts)
    got_c, got_p = cells.cpu().numpy().tobytes(), cprf.cpu().numpy().tobytes()
    mod.coalesce_enable(ts, False)  # every call below runs alone: the n = 1 path
    try:
        for i, blob in enumerate(blobs):
            c, p = mod.compute_cells_and_kzg_proofs(blob, ts)
            assert b"".join(c) == got_c[262144 * i : 262144 * (i + 1)], i
            assert b"".join(p) == got_p[6144 * i : 6144 * (i + 1)], i
    finally:
        mod.coalesce_enable(ts, True)


def test_device_mode_ordering_against_the_callers_streams(env):
    """DEVICE-mode contract (include/ckzg_b200.h): engine calls run on blocking streams, so writes enqueued on the legacy
    default stream are seen without a synchronize; a producer on a non-blocking side stream is named with
    ckzg_b200_set_caller_stream.  Here the CORRECT proofs are written into a buffer of wrong ones behind ~20 ms of
    queued work, and the call must see them."""
    import torch

    mod, ts, n, host, dev, cms, prs = env
    m = 256
    wrong = prs[: 48 * m].roll(48)  # every proof belongs to the neighbouring blob
    filler = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    # (1) legacy default stream, no synchronize
    buf = wrong.clone()
    for _ in range(8):
        filler.fill_(3)
    buf.copy_(prs[: 48 * m])
    assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), buf.data_ptr(), m, ts) is True
    # (2) a non-blocking side stream, named for this thread
    side = torch.cuda.Stream()
    buf2 = wrong.clone()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(8):
            filler.fill_(5)
        buf2.copy_(prs[: 48 * m])
    mod.set_caller_stream(side.cuda_stream)
    try:
        assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), buf2.data_ptr(), m, ts) is True
    finally:
        mod.set_caller_stream(0)
    torch.cuda.synchronize()
    assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), wrong.data_ptr(), m, ts) is False


@pytest.mark.parametrize("pieces", ["0", "8"])
def test_pinned_host_batch_with_and_without_column_pieces(pieces):
    """Host-pointer batches from PINNED memory, with the default arrangement (one hash launch per 512-blob chunk) and
    with the experimental one (CKZG_B200_TAIL_PIECES=8: the tail sent in eight column pieces and hashed piece by piece
    with the SHA state carried between launches, verify.cu launch_blob_challenges_range).  A wrong state hand-over
    would change z and with it the verdict.  The switch is read once per process, hence the subprocess
    (tests/pieces_check.py)."""
    import subprocess

    env2 = dict(os.environ, CKZG_B200_TAIL_PIECES=pieces)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "pieces_check.py")], env=env2, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "pieces_check ok" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
