"""Argument checks of the frozen C API that the reference performs before any arithmetic, exercised WITHOUT a GPU:
they must answer the same C_KZG_RET as the reference (src/setup/setup.c:411-430, src/eip7594/eip7594.c:72-74,
191-213, 852-864; src/eip4844/eip4844.c:791-794) before the engine is ever asked to do anything.  The same calls
are made against the unmodified reference build (oracle/_ref) to pin the expected codes."""
import ctypes as C
import os

import pytest

import __graft_entry__ as entry
from oracle import ref_lib

OK, BADARGS, ERROR = 0, 1, 2


def libs():
    mod = entry.load_package()
    if not os.path.exists(mod.LIB_PATH):
        pytest.skip("libckzg_b200.so not built")
    out = [("b200", C.CDLL(mod.LIB_PATH))]
    if os.path.exists(ref_lib.REF_SO):
        out.append(("reference", C.CDLL(ref_lib.REF_SO)))
    return out


def setup_bytes():
    mod = entry.load_package()
    toks = open(mod.SETUP_TXT).read().split()
    n1, n2 = int(toks[0]), int(toks[1])
    lag = bytes.fromhex("".join(toks[2 : 2 + n1]))
    g2 = bytes.fromhex("".join(toks[2 + n1 : 2 + n1 + n2]))
    mono = bytes.fromhex("".join(toks[2 + n1 + n2 : 2 + 2 * n1 + n2]))
    return mono, lag, g2


def test_load_trusted_setup_rejects_bad_sizes_and_precompute_before_touching_a_device():
    mono, lag, g2 = setup_bytes()
    for name, lib in libs():
        s = C.create_string_buffer(80)
        call = lambda m, l, g, pre: lib.load_trusted_setup(s, m, C.c_uint64(len(m)), l, C.c_uint64(len(l)), g, C.c_uint64(len(g)), C.c_uint64(pre))
        assert call(mono, lag, g2, 16) == BADARGS, name  # setup.c:411
        assert call(mono[:-48], lag, g2, 0) == BADARGS, name  # setup.c:425-430
        assert call(mono, lag + bytes(48), g2, 0) == BADARGS, name
        assert call(mono, lag, g2[:-96], 0) == BADARGS, name
        assert s.raw == bytes(80), name  # the struct is left in the state free_trusted_setup accepts
        lib.free_trusted_setup.restype = None
        lib.free_trusted_setup(s)
        lib.free_trusted_setup(s)
        lib.free_trusted_setup(None)  # setup.c:162-190: NULL is fine


def test_index_and_count_checks_come_first():
    """With a zeroed KZGSettings (no engine behind it) the reference-ordered checks still decide the result."""
    name, lib = libs()[0]
    s = C.create_string_buffer(80)
    ok = C.c_bool(True)
    cells = bytes(2048 * 64)
    out_c, out_p = C.create_string_buffer(128 * 2048), C.create_string_buffer(128 * 48)
    U64 = C.c_uint64
    # verify_cell_kzg_proof_batch: n = 0 is valid (eip7594.c:852-855), an index >= 128 is BADARGS (:861-864)
    assert lib.verify_cell_kzg_proof_batch(C.byref(ok), None, None, None, None, U64(0), s) == OK and ok.value is True
    idx = (U64 * 2)(0, 128)
    assert lib.verify_cell_kzg_proof_batch(C.byref(ok), bytes(96), idx, bytes(4096), bytes(96), U64(2), s) == BADARGS and ok.value is False
    # verify_blob_kzg_proof_batch: n = 0 is valid (eip4844.c:791-794)
    ok.value = False
    assert lib.verify_blob_kzg_proof_batch(C.byref(ok), None, None, None, U64(0), s) == OK and ok.value is True
    # compute_cells_and_kzg_proofs: both outputs NULL (eip7594.c:72-74)
    assert lib.compute_cells_and_kzg_proofs(None, None, bytes(131072), s) == BADARGS
    # recover_cells_and_kzg_proofs: counts and indices (eip7594.c:191-213)
    asc = (U64 * 64)(*range(64))
    assert lib.recover_cells_and_kzg_proofs(out_c, out_p, asc, cells, U64(63), s) == BADARGS  # fewer than half
    assert lib.recover_cells_and_kzg_proofs(out_c, out_p, asc, cells, U64(129), s) == BADARGS  # more than all
    dup = (U64 * 64)(*([0, 0] + list(range(2, 64))))
    assert lib.recover_cells_and_kzg_proofs(out_c, out_p, dup, cells, U64(64), s) == BADARGS  # not strictly ascending
    big = (U64 * 64)(*(list(range(63)) + [128]))
    assert lib.recover_cells_and_kzg_proofs(out_c, out_p, big, cells, U64(64), s) == BADARGS  # index out of range
