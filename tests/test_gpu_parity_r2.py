"""-m gpu: the parity gaps the round-1 review listed (VERDICT "What's weak" 1a-1f), each closed against the
UNMODIFIED reference build (oracle/_ref) on the same inputs:

  a. the BATCHED FK20 path (msm_affine.cu, default for >= 8 blobs) and batched recovery, sampled blob by blob
     against the reference (before: only compared with the engine's own single-blob path);
  b. verify_cell_kzg_proof_batch at 256 blobs x 128 cells, the true case and every negative control mirrored
     on the reference;
  c. `precompute` = 8 passed to the product (BASELINE configs[2]/[3]) on the cells vectors;
  d. verify_blob_kzg_proof_batch at n = 4096 (the headline size) -- true / one bad proof / one non-canonical
     field element -- with the reference's verdict and return code;
  e. the coalesced per-blob API against the reference (before: against the batched engine output);
  f. load_trusted_setup error branches on the device (src/setup/setup.c:339-358, :447-477).
"""
import ctypes as C
import os
import random
import sys
import threading

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BLOB = 131072
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def _special_blobs():
    z, one, top = (0).to_bytes(32, "big"), (1).to_bytes(32, "big"), (R - 1).to_bytes(32, "big")
    return [
        z * 4096,
        one * 4096,
        top * 4096,
        z * 4000 + top + z * 95,
        b"".join(i.to_bytes(32, "big") for i in range(4096)),
        (one + top) * 2048,
        z * 64 + one * 64 + z * 3968,
        b"".join(((1 << 254) + 7 * i).to_bytes(32, "big") for i in range(4096)),
    ]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_lib

    if not os.path.exists(ref_lib.REF_SO):
        pytest.skip("oracle/_ref not built")
    k = ref_lib.CKZG(precompute=8)  # same outputs as precompute = 0 (only the reference's speed differs)
    yield k
    k.close()


@pytest.fixture(scope="module")
def env():
    """256 blobs (8 structured + 248 random), their cells / FK20 proofs from ONE batched engine call."""
    import torch

    import __graft_entry__ as entry
    import bench

    mod = entry.load_package()
    ts = mod.load_trusted_setup()
    n = 256
    special = _special_blobs()
    rnd = bench.synth_blobs(n - len(special), 20262).tobytes()
    hb = b"".join(special) + rnd
    dev = torch.frombuffer(bytearray(hb), dtype=torch.uint8).cuda()
    cells = torch.empty(n * 2 * BLOB, dtype=torch.uint8, device="cuda")
    cprf = torch.empty(n * 128 * 48, dtype=torch.uint8, device="cuda")
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), cprf.data_ptr(), dev.data_ptr(), n, ts)
    cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
    return dict(mod=mod, ts=ts, n=n, hb=hb, dev=dev, cells=cells, cprf=cprf, cms=cms)


SAMPLE = list(range(8)) + random.Random(12).sample(range(8, 256), 6)


def test_batched_cells_and_proofs_vs_reference_sampled(env, ref):
    """(a) 14 of the 256 blobs of one batched call -- all eight structured blobs and six random ones."""
    hc, hp = env["cells"].cpu().numpy().tobytes(), env["cprf"].cpu().numpy().tobytes()
    for i in SAMPLE:
        want_c, want_p = ref.compute_cells_and_kzg_proofs(env["hb"][BLOB * i : BLOB * (i + 1)])
        assert hc[2 * BLOB * i : 2 * BLOB * (i + 1)] == want_c, i
        assert hp[6144 * i : 6144 * (i + 1)] == want_p, i


def test_batched_recovery_vs_reference_sampled(env, ref):
    """(a) batched recovery (three erasure patterns) against the reference's recovery of the same cells."""
    import torch

    mod, ts, n = env["mod"], env["ts"], env["n"]
    cells = env["cells"]
    hc = cells.cpu().numpy().tobytes()
    patterns = (list(range(0, 128, 2)), list(range(64, 128)), sorted(random.Random(3).sample(range(128), 90)))
    for pi, pattern in enumerate(patterns):
        given = cells.view(n, 128, 2048)[:, pattern, :].contiguous()
        rc = torch.empty_like(cells)
        rp = torch.empty_like(env["cprf"])
        mod.recover_cells_and_kzg_proofs_device(rc.data_ptr(), rp.data_ptr(), pattern * n, given.data_ptr(), len(pattern), n, ts)
        got_c, got_p = rc.cpu().numpy().tobytes(), rp.cpu().numpy().tobytes()
        for i in (SAMPLE[pi::3] + [pi]):  # a third of the sample per pattern (the reference takes 0.25 s per blob)
            sub = b"".join(hc[(i * 128 + k) * 2048 : (i * 128 + k + 1) * 2048] for k in pattern)
            want_c, want_p = ref.recover_cells_and_kzg_proofs(pattern, sub)
            assert got_c[2 * BLOB * i : 2 * BLOB * (i + 1)] == want_c, (pi, i)
            assert got_p[6144 * i : 6144 * (i + 1)] == want_p, (pi, i)


def test_verify_cells_256x128_and_controls_mirrored_on_reference(env, ref):
    """(b) 32,768 cells in one call: true case, and four corruptions; the reference gives its verdict on the same
    full batch for the true case and on the 8-blob sub-batch that contains the corruption for the controls (a batch
    is valid iff its sub-batches are; the full batch costs the reference 6 s per verdict)."""
    import torch

    mod, ts, n = env["mod"], env["ts"], env["n"]
    h_cells = env["cells"].cpu()
    h_cprf = env["cprf"].cpu()
    h_cm_rows = env["cms"].cpu().view(-1, 48).repeat_interleave(128, dim=0).contiguous()
    idx = [k for _ in range(n) for k in range(128)]

    def engine(cm_rows, idx_l, cells_t, prf_t, count=n * 128, first=0):
        arr = (C.c_uint64 * count)(*idx_l[first : first + count])
        return mod.verify_cell_kzg_proof_batch_ptr(
            cm_rows.data_ptr() + 48 * first, arr, cells_t.data_ptr() + 2048 * first, prf_t.data_ptr() + 48 * first, count, ts
        )

    def reference(cm_rows, idx_l, cells_t, prf_t, count, first):
        return ref.verify_cell_kzg_proof_batch(
            cm_rows.numpy().tobytes()[48 * first : 48 * (first + count)], idx_l[first : first + count],
            cells_t.numpy().tobytes()[2048 * first : 2048 * (first + count)], prf_t.numpy().tobytes()[48 * first : 48 * (first + count)],
        )

    assert engine(h_cm_rows, idx, h_cells, h_cprf) is True
    assert reference(h_cm_rows, idx, h_cells, h_cprf, n * 128, 0) is True
    sub = 8 * 128  # the sub-batch given to the reference: blobs [8j, 8j + 8)

    def control(make):
        cm2, idx2, cells2, prf2 = h_cm_rows.clone(), list(idx), h_cells.clone(), h_cprf.clone()
        where = make(cm2, idx2, cells2, prf2)  # tuple index of the corruption
        first = (where // sub) * sub
        got_full = engine(cm2, idx2, cells2, prf2)
        got_sub = engine(cm2, idx2, cells2, prf2, sub, first)
        want_sub = reference(cm2, idx2, cells2, prf2, sub, first)
        assert got_full is False and got_sub is False and want_sub is False, (where, got_full, got_sub, want_sub)

    def swap_proofs(cm2, idx2, cells2, prf2):
        a, b = 20001, 20002
        t = prf2[48 * a : 48 * a + 48].clone()
        prf2[48 * a : 48 * a + 48] = prf2[48 * b : 48 * b + 48]
        prf2[48 * b : 48 * b + 48] = t
        return a

    def flip_cell_bit(cm2, idx2, cells2, prf2):
        a = 31000
        cells2[2048 * a + 2047] ^= 1
        return a

    def wrong_index(cm2, idx2, cells2, prf2):
        a = 128 * 50 + 77  # a random blob (blob 0 is the zero polynomial: every index verifies, also in the reference)
        idx2[a] = (idx2[a] + 64) % 128
        return a

    def wrong_commitment(cm2, idx2, cells2, prf2):
        a = 128 * 100 + 3
        cm2[a] = h_cm_rows[0]
        return a

    for make in (swap_proofs, flip_cell_bit, wrong_index, wrong_commitment):
        control(make)
    # invalid encodings: BADARGS from both
    from oracle import ref_lib

    cells3 = h_cells.clone()
    cells3[2048 * 5000 + 64 : 2048 * 5000 + 96] = torch.frombuffer(bytearray(R.to_bytes(32, "big")), dtype=torch.uint8)
    with pytest.raises(ValueError):
        engine(h_cm_rows, idx, cells3, h_cprf)
    with pytest.raises(ref_lib.BadArgs):
        reference(h_cm_rows, idx, cells3, h_cprf, sub, (5000 // sub) * sub)


def test_precompute_8_on_the_cells_vectors():
    """(c) BASELINE configs[2]/[3] say precompute = 8: the product accepts it (range-checked as setup.c:411, the
    table width itself follows the HBM plan) and reproduces the consensus vectors and the default setup's bytes."""
    import vector_runner as vr
    from gpu_common import product, synth_blob

    k8 = product(precompute=8)
    k0 = product(precompute=0)
    try:
        for api in ("compute_cells_and_kzg_proofs", "recover_cells_and_kzg_proofs", "compute_cells"):
            bad, cnt = vr.run_api(api, k8)
            assert cnt > 0 and not bad, (api, [b[0] for b in bad][:3])
        blob = synth_blob(8)
        assert k8.compute_cells_and_kzg_proofs(blob) == k0.compute_cells_and_kzg_proofs(blob)
        assert k8.blob_to_kzg_commitment(blob) == k0.blob_to_kzg_commitment(blob)
        from oracle import ref_lib

        with pytest.raises(ref_lib.BadArgs):
            product(precompute=16)  # setup.c:411
    finally:
        k8.close()
        k0.close()


def test_verify_blob_batch_n4096_with_reference_verdicts(ref):
    """(d) the headline configuration as a pytest: 4096 blobs, commitments and proofs from the engine (12 of them
    re-derived by the reference), then true / one bad proof / one non-canonical element with the reference's
    verdict and return code on the same 4096-blob inputs (5 s per reference call)."""
    import torch

    import __graft_entry__ as entry
    import bench
    from oracle import ref_lib

    mod = entry.load_package()
    ts = mod.load_trusted_setup()
    n = 4096
    host = torch.from_numpy(bench.synth_blobs(n, 4096_2))
    dev = host.cuda()
    cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
    mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
    hb, hc, hp = host.numpy().tobytes(), cms.cpu().numpy().tobytes(), prs.cpu().numpy().tobytes()
    for i in random.Random(40).sample(range(n), 10) + [0, n - 1]:
        c = ref.blob_to_kzg_commitment(hb[BLOB * i : BLOB * (i + 1)])
        assert hc[48 * i : 48 * i + 48] == c
        assert hp[48 * i : 48 * i + 48] == ref.compute_blob_kzg_proof(hb[BLOB * i : BLOB * (i + 1)], c)
    # true
    assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts) is True
    assert mod.verify_blob_kzg_proof_batch(hb, hc, hp, ts) is True  # frozen API, host bytes (pageable)
    assert ref.verify_blob_kzg_proof_batch(hb, hc, hp) is True
    # one bad proof (a valid point: proof of another blob)
    i = 2901
    hp2 = hp[: 48 * i] + hp[48 * (i + 1) : 48 * (i + 2)] + hp[48 * (i + 1) :]
    assert len(hp2) == len(hp)
    assert mod.verify_blob_kzg_proof_batch(hb, hc, hp2, ts) is False
    bad = torch.frombuffer(bytearray(hp2), dtype=torch.uint8).cuda()
    assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), bad.data_ptr(), n, ts) is False
    assert ref.verify_blob_kzg_proof_batch(hb, hc, hp2) is False
    # one non-canonical field element deep inside the batch -> BADARGS from both
    j = 3333
    hb3 = hb[: BLOB * j + 32 * 4000] + R.to_bytes(32, "big") + hb[BLOB * j + 32 * 4001 :]
    with pytest.raises(ValueError):
        mod.verify_blob_kzg_proof_batch(hb3, hc, hp, ts)
    with pytest.raises(ref_lib.BadArgs):
        ref.verify_blob_kzg_proof_batch(hb3, hc, hp)
    ts.close()


def test_coalesced_per_blob_api_vs_reference(env, ref):
    """(e) 16 threads on the frozen per-blob API at once (merged into batched engine calls by the coalescer):
    every caller's bytes are the reference's bytes for its blob."""
    mod, ts, hb = env["mod"], env["ts"], env["hb"]
    picks = SAMPLE + [100, 200]
    want = {}
    for i in picks:
        blob = hb[BLOB * i : BLOB * (i + 1)]
        c = ref.blob_to_kzg_commitment(blob)
        want[i] = (c, ref.compute_blob_kzg_proof(blob, c), ref.compute_cells_and_kzg_proofs(blob))
    errors = []
    before = mod.coalesce_stats(ts)

    def worker(i):
        try:
            blob = hb[BLOB * i : BLOB * (i + 1)]
            c, p, (cells, proofs) = want[i]
            for _ in range(2):
                assert mod.blob_to_kzg_commitment(blob, ts) == c
                assert mod.compute_blob_kzg_proof(blob, c, ts) == p
                gc, gp = mod.compute_cells_and_kzg_proofs(blob, ts)
                assert b"".join(gc) == cells and b"".join(gp) == proofs
                idx = list(range(i % 2, 128, 2))
                rc, rp = mod.recover_cells_and_kzg_proofs(idx, [gc[k] for k in idx], ts)
                assert b"".join(rc) == cells and b"".join(rp) == proofs
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)[:200]))

    threads = [threading.Thread(target=worker, args=(i,)) for i in picks]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    after = mod.coalesce_stats(ts)
    assert after["compute_cells_and_kzg_proofs"][1] - before["compute_cells_and_kzg_proofs"][1] < 2 * len(picks)  # calls were merged


def test_load_trusted_setup_error_branches_on_device(ref):
    """(f) the BADARGS branches of load_trusted_setup run on the GPU here (monomial_form_kernel, g1/g2 uncompress):
    same return code as the reference for a monomial-form setup in the Lagrange slot (setup.c:339-358), a G1 point
    off the curve, a G1 x-coordinate >= p, an uncompressed-flagged G1 point and G2 points pushed off the curve (setup.c:447-477);
    G2 encodings that decode to a DIFFERENT curve point load fine in the reference, and here."""
    from gpu_common import product_lib_path
    from oracle import ref_lib

    mono, lag, g2 = ref_lib.parse_trusted_setup_text(ref_lib.SETUP_TXT)
    libs = {"ref": C.CDLL(ref_lib.REF_SO), "gpu": C.CDLL(product_lib_path())}

    def load(lib, m, l, g, pre=0):
        s = C.create_string_buffer(80)
        fn = lib.load_trusted_setup
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_uint64]
        rc = fn(s, m, len(m), l, len(l), g, len(g), pre)
        lib.free_trusted_setup.restype = None
        lib.free_trusted_setup(s)
        lib.free_trusted_setup(s)  # twice: allowed after a failed load (setup.c:365-376)
        return rc

    def find_off_curve(b48):
        """flip low bits of x until x^3 + 4 is a non-residue (the reference rejects in blst_p1_uncompress)"""
        P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
        x0 = int.from_bytes(b48, "big") & ((1 << 381) - 1)
        for d in range(1, 64):
            x = x0 ^ d
            if x < P and pow((x * x * x + 4) % P, (P - 1) // 2, P) != 1:
                return ((int.from_bytes(b48, "big") >> 381 << 381) | x).to_bytes(48, "big")
        raise AssertionError

    cases = {
        "valid": (mono, lag, g2),
        "monomial_in_lagrange_slot": (mono, mono, g2),
        "lagrange_in_both_slots": (lag, lag, g2),  # accepted by the reference's check (only the Lagrange slot is tested)
        "g1_monomial_off_curve": (mono[: 48 * 7] + find_off_curve(mono[48 * 7 : 48 * 8]) + mono[48 * 8 :], lag, g2),
        "g1_lagrange_off_curve": (mono, lag[: 48 * 4095] + find_off_curve(lag[48 * 4095 :]), g2),
        "g1_x_not_below_p": (mono, lag[:48] + bytes([0x9F]) + b"\xff" * 47 + lag[96:], g2),
        "g1_uncompressed_flag": (mono[:48] + bytes([mono[48] & 0x7F]) + mono[49:], lag, g2),
        # last byte of a G2 x-coordinate, low bits flipped: d = 3 / 2 / 1 leave the curve (BADARGS); d = 1 / 1 / 2 land on
        # another curve point, which the reference accepts (no subgroup check at setup, setup.c:468-477) -- so must we
        "g2_0_off_curve": (mono, lag, g2[:95] + bytes([g2[95] ^ 3]) + g2[96:]),
        "g2_1_off_curve": (mono, lag, g2[: 96 + 95] + bytes([g2[96 + 95] ^ 2]) + g2[96 * 2 :]),
        "g2_64_off_curve": (mono, lag, g2[: 96 * 64 + 95] + bytes([g2[96 * 64 + 95] ^ 1])),
        "g2_0_other_point": (mono, lag, g2[:95] + bytes([g2[95] ^ 1]) + g2[96:]),
        "g2_64_other_point": (mono, lag, g2[: 96 * 64 + 95] + bytes([g2[96 * 64 + 95] ^ 2])),
        "short_g1": (mono[:-48], lag, g2),
    }
    for name, (m, l, g) in cases.items():
        want = load(libs["ref"], m, l, g)
        got = load(libs["gpu"], m, l, g)
        assert got == want, (name, got, want)
        accepted = name in ("valid", "lagrange_in_both_slots", "g2_0_other_point", "g2_64_other_point")
        assert want == (0 if accepted else 1), (name, want)
    assert load(libs["gpu"], mono, lag, g2, 16) == load(libs["ref"], mono, lag, g2, 16) == 1


def test_g1_fft_one_thread_per_butterfly_matches_quad_form_and_reference(env, ref, monkeypatch):
    """Batches of >= 1024 blobs run the FK20 G1 FFT stages with one thread per butterfly (fk20_fft.cu
    g1_fft_stage_thread_kernel), smaller ones with four lanes per group operation.  Forced on for the 256-blob batch of
    this module (8 structured blobs included): identical bytes to the quad form, which the tests above pin against the
    reference -- and six blobs are compared with the reference directly."""
    import torch

    mod, ts, n = env["mod"], env["ts"], env["n"]
    monkeypatch.setenv("CKZG_B200_FFT_THREAD_MIN", "1")
    cells = torch.empty_like(env["cells"])
    cprf = torch.empty_like(env["cprf"])
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), cprf.data_ptr(), env["dev"].data_ptr(), n, ts)
    monkeypatch.delenv("CKZG_B200_FFT_THREAD_MIN")
    assert torch.equal(cells, env["cells"])
    assert torch.equal(cprf, env["cprf"])
    hp = cprf.cpu().numpy().tobytes()
    for i in SAMPLE[::3] + [255]:
        _, want_p = ref.compute_cells_and_kzg_proofs(env["hb"][BLOB * i : BLOB * (i + 1)], want_cells=False)
        assert hp[6144 * i : 6144 * (i + 1)] == want_p, i
