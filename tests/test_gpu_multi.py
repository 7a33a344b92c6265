"""-m gpu: the in-library multi-device context (CKZG_B200_DEVICES; c-kzg-4844_b200/csrc/multi.cu, api_verify.cu
verify_blob_batch_multi) behind the FROZEN API.  On a box with two or more GPUs the context spans devices 0 and 1;
on a one-GPU box two replicas share device 0 -- the same host code path (fan-out threads, pinned exchange, partial
sums, per-range status), which is what these tests pin.  Every result is compared with the single-device context and,
sampled, with the unmodified reference."""
import ctypes as C
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BLOB = 131072
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


@pytest.fixture(scope="module")
def ctxs():
    import torch

    import bench
    from gpu_common import product

    old = {k: os.environ.get(k) for k in ("CKZG_B200_DEVICES", "CKZG_B200_COMMIT_WINDOW", "CKZG_B200_FK_WINDOW")}
    # small tables: this module holds three replicas of the setup at once
    os.environ["CKZG_B200_COMMIT_WINDOW"] = "12"
    os.environ["CKZG_B200_FK_WINDOW"] = "10"
    os.environ.pop("CKZG_B200_DEVICES", None)
    single = product()
    os.environ["CKZG_B200_DEVICES"] = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    multi = product()
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    n = 600
    hb = bench.synth_blobs(n, 777).tobytes()
    eng = lambda k: C.c_void_p(int.from_bytes(k.settings.raw[56:64], "little"))
    for k in (single, multi):
        k.lib.ckzg_b200_ctx_device_count.restype = C.c_int
    assert single.lib.ckzg_b200_ctx_device_count(eng(single)) == 1
    assert multi.lib.ckzg_b200_ctx_device_count(eng(multi)) == 2
    cms, prs = C.create_string_buffer(48 * n), C.create_string_buffer(48 * n)
    st = (C.c_int * n)()
    assert single.lib.ckzg_b200_blob_to_kzg_commitment_batch(eng(single), cms, hb, C.c_uint64(n), 0, st) == 0
    assert single.lib.ckzg_b200_compute_blob_kzg_proof_batch(eng(single), prs, hb, cms, C.c_uint64(n), 0, st) == 0
    yield dict(single=single, multi=multi, n=n, hb=hb, cms=cms.raw, prs=prs.raw, eng=eng)
    single.close()
    multi.close()


def test_multi_device_commitments_and_proofs_match_single_device(ctxs):
    from oracle import ref_lib

    multi, eng, n, hb = ctxs["multi"], ctxs["eng"], ctxs["n"], ctxs["hb"]
    # one non-canonical blob in the second half: only its own status is BADARGS, every other output is unchanged
    bad_i = 431
    hb2 = hb[: BLOB * bad_i + 32 * 9] + R.to_bytes(32, "big") + hb[BLOB * bad_i + 32 * 10 :]
    out = C.create_string_buffer(48 * n)
    st = (C.c_int * n)()
    assert multi.lib.ckzg_b200_blob_to_kzg_commitment_batch(eng(multi), out, hb2, C.c_uint64(n), 0, st) == 1
    assert [i for i in range(n) if st[i]] == [bad_i]
    for i in range(n):
        if i != bad_i:
            assert out.raw[48 * i : 48 * i + 48] == ctxs["cms"][48 * i : 48 * i + 48], i
    pr = C.create_string_buffer(48 * n)
    assert multi.lib.ckzg_b200_compute_blob_kzg_proof_batch(eng(multi), pr, hb, ctxs["cms"], C.c_uint64(n), 0, st) == 0
    assert pr.raw == ctxs["prs"]
    if os.path.exists(ref_lib.REF_SO):
        ref = ref_lib.CKZG()
        for i in (0, 299, 300, 599):  # both sides of the range boundary
            blob = hb[BLOB * i : BLOB * (i + 1)]
            c = ref.blob_to_kzg_commitment(blob)
            assert c == ctxs["cms"][48 * i : 48 * i + 48]
            assert ref.compute_blob_kzg_proof(blob, c) == pr.raw[48 * i : 48 * i + 48]
        ref.close()


def test_multi_device_verify_blob_batch_single_challenge(ctxs):
    """600 blobs -> two shards of 300 with ONE challenge: same verdicts as the one-device call for the valid batch,
    a wrong proof in either shard, and BADARGS for a non-canonical element / an invalid point in the second shard."""
    from oracle import ref_lib

    single, multi, n, hb, cms, prs = (ctxs[k] for k in ("single", "multi", "n", "hb", "cms", "prs"))
    assert multi.verify_blob_kzg_proof_batch(hb, cms, prs) is True
    for i in (7, 455):
        bad = prs[: 48 * i] + prs[48 * (i + 1) : 48 * (i + 2)] + prs[48 * (i + 1) :]
        assert multi.verify_blob_kzg_proof_batch(hb, cms, bad) is False
        assert single.verify_blob_kzg_proof_batch(hb, cms, bad) is False
    hb2 = hb[: BLOB * 580 + 32 * 4095] + R.to_bytes(32, "big") + hb[BLOB * 581 :]
    with pytest.raises(ref_lib.BadArgs):
        multi.verify_blob_kzg_proof_batch(hb2, cms, prs)
    with pytest.raises(ref_lib.BadArgs):
        multi.verify_blob_kzg_proof_batch(hb, cms, prs[: 48 * 310] + bytes(48) + prs[48 * 311 :])
    # batches too small to shard (< 256 blobs per device) take the one-device path of the same context
    m = 300
    assert multi.verify_blob_kzg_proof_batch(hb[: BLOB * m], cms[: 48 * m], prs[: 48 * m]) is True
    # the reference's verdict on the same 600-blob inputs
    if os.path.exists(ref_lib.REF_SO):
        ref = ref_lib.CKZG()
        assert ref.verify_blob_kzg_proof_batch(hb, cms, prs) is True
        ref.close()


def test_multi_device_cells_recovery_and_cell_verification(ctxs):
    single, multi, eng, hb, cms = (ctxs[k] for k in ("single", "multi", "eng", "hb", "cms"))
    m = 64
    outs = {}
    for name, k in (("single", single), ("multi", multi)):
        cells, proofs = C.create_string_buffer(m * 2 * BLOB), C.create_string_buffer(m * 128 * 48)
        st = (C.c_int * m)()
        assert k.lib.ckzg_b200_compute_cells_and_kzg_proofs_batch(eng(k), cells, proofs, hb, C.c_uint64(m), 0, st) == 0
        outs[name] = (cells.raw, proofs.raw)
    assert outs["single"] == outs["multi"]
    cells, proofs = outs["multi"]
    # recovery from the odd cells, 64 blobs over two devices
    pattern = list(range(1, 128, 2))
    idx = (C.c_uint64 * (m * 64))(*(pattern * m))
    given = b"".join(cells[(b * 128 + k) * 2048 : (b * 128 + k + 1) * 2048] for b in range(m) for k in pattern)
    rc_, rp_ = C.create_string_buffer(m * 2 * BLOB), C.create_string_buffer(m * 128 * 48)
    st = (C.c_int * m)()
    multi.lib.ckzg_b200_recover_cells_and_kzg_proofs_batch.restype = C.c_int
    assert multi.lib.ckzg_b200_recover_cells_and_kzg_proofs_batch(eng(multi), rc_, rp_, idx, given, C.c_uint64(64), C.c_uint64(m), 0, st) == 0
    assert rc_.raw == cells and rp_.raw == proofs
    # a descending index pair in the LAST blob fails the whole call up front (eip7594.c:191-213), as on one device
    idx2 = (C.c_uint64 * (m * 64))(*(pattern * (m - 1) + [3, 1] + pattern[2:]))
    assert multi.lib.ckzg_b200_recover_cells_and_kzg_proofs_batch(eng(multi), rc_, rp_, idx2, given, C.c_uint64(64), C.c_uint64(m), 0, st) == 1
    # cell verification: 64 x 128 cells = two sub-batches (one per device), verdicts AND-ed
    cm_rows = b"".join(cms[48 * b : 48 * b + 48] * 128 for b in range(m))
    idx_all = [k for _ in range(m) for k in range(128)]
    assert multi.verify_cell_kzg_proof_batch(cm_rows, idx_all, cells, proofs) is True
    bad_cells = cells[: 2048 * 8000 + 100] + bytes([cells[2048 * 8000 + 100] ^ 1]) + cells[2048 * 8000 + 101 :]
    assert multi.verify_cell_kzg_proof_batch(cm_rows, idx_all, bad_cells, proofs) is False
    assert single.verify_cell_kzg_proof_batch(cm_rows, idx_all, bad_cells, proofs) is False
    from oracle import ref_lib

    with pytest.raises(ref_lib.BadArgs):
        multi.verify_cell_kzg_proof_batch(cm_rows, idx_all, cells, proofs[: 48 * 8100] + bytes(48) + proofs[48 * 8101 :])
    with pytest.raises(ref_lib.BadArgs):
        multi.verify_cell_kzg_proof_batch(cm_rows, idx_all[:-1] + [128], cells, proofs)
