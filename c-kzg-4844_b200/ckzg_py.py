"""Host-side mirror of the reference Python binding (bindings/python/ckzg_wrap.c) over ctypes.

Same names, argument order and error behaviour as the `ckzg` module of the reference:
    ts = load_trusted_setup(path, precompute)
    blob_to_kzg_commitment(blob, ts) -> bytes48
    compute_kzg_proof(blob, z, ts) -> (proof48, y32)
    compute_blob_kzg_proof(blob, commitment, ts) -> bytes48
    verify_kzg_proof(commitment, z, y, proof, ts) -> bool
    verify_blob_kzg_proof(blob, commitment, proof, ts) -> bool
    verify_blob_kzg_proof_batch(blobs, commitments, proofs, ts) -> bool     (concatenated bytes)
    compute_cells_and_kzg_proofs(blob, ts) -> (cells[128], proofs[128])
    recover_cells_and_kzg_proofs(cell_indices, cells, ts) -> (cells[128], proofs[128])
    verify_cell_kzg_proof_batch(commitments, cell_indices, cells, proofs, ts) -> bool
Invalid input raises ValueError (the binding's mapping of C_KZG_BADARGS, ckzg_wrap.c), anything else
RuntimeError.  Plus the engine's batched / device-pointer entry points (include/ckzg_b200.h).

No arithmetic happens here and nothing under oracle/ is imported: the library either runs on the
GPU or raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CKZG_B200_LIB") or os.path.join(_HERE, "libckzg_b200.so")  # the override: A/B builds of the same library
SETUP_TXT = os.path.join(_HERE, "data", "trusted_setup.txt")

BYTES_PER_BLOB = 131072
BYTES_PER_CELL = 2048
CELLS_PER_EXT_BLOB = 128
HOST, DEVICE = 0, 1

_lib = None


def lib():
    """The product shared library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libckzg_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        for name in (
            "load_trusted_setup blob_to_kzg_commitment compute_kzg_proof compute_blob_kzg_proof verify_kzg_proof "
            "verify_blob_kzg_proof verify_blob_kzg_proof_batch ckzg_b200_blob_to_kzg_commitment_batch "
            "ckzg_b200_compute_blob_kzg_proof_batch ckzg_b200_compute_kzg_proof_batch ckzg_b200_verify_blob_kzg_proof_batch "
            "ckzg_b200_verify_kzg_proof ckzg_b200_verify_shard_stage1 ckzg_b200_verify_shard_stage2 "
            "ckzg_b200_verify_shard_finish ckzg_b200_pack_verify_tuples ckzg_b200_compute_challenge ckzg_b200_selftest_field ckzg_b200_selftest_g1"
        ).split():
            getattr(_lib, name).restype = C.c_int
        _lib.ckzg_b200_launch_count.restype = C.c_uint64
        _lib.ckzg_b200_launch_count.argtypes = [C.c_void_p]
        _lib.free_trusted_setup.restype = None
        _lib.ckzg_b200_verify_shard_free.restype = None
        _lib.ckzg_b200_verify_shard_free.argtypes = [C.c_void_p]
        _lib.ckzg_b200_set_caller_stream.restype = None
        _lib.ckzg_b200_set_caller_stream.argtypes = [C.c_void_p]
    return _lib


def _raise(code, fn):
    """C_KZG_RET -> exception; the code rides along as `.code` (parallel.py agrees on it across ranks)."""
    if code == 0:
        return
    if code == 1:
        e = ValueError("%s: invalid argument (C_KZG_BADARGS)" % fn)
    elif code == 3:
        e = MemoryError("%s: C_KZG_MALLOC" % fn)
    else:
        e = RuntimeError("%s: C_KZG_ERROR (code %d) -- CUDA device/engine failure; there is no CPU path" % (fn, code))
    e.code = code
    raise e


class TrustedSetup:
    """Owns the caller-allocated 80-byte KZGSettings (src/setup/settings.h:27-79)."""

    def __init__(self, path=SETUP_TXT, precompute=0):
        self._s = C.create_string_buffer(80)
        with open(path) as f:
            tok = f.read().split()
        n1, n2 = int(tok[0]), int(tok[1])
        body = tok[2:]
        lag = bytes.fromhex("".join(body[:n1]))
        g2 = bytes.fromhex("".join(body[n1 : n1 + n2]))
        mono = bytes.fromhex("".join(body[n1 + n2 : n1 + n2 + n1]))
        fn = lib().load_trusted_setup
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_uint64]
        _raise(fn(self._s, mono, len(mono), lag, len(lag), g2, len(g2), precompute), "load_trusted_setup")
        self.precompute = precompute

    @property
    def ptr(self):
        return self._s

    @property
    def engine(self):
        """The ckzg_b200_ctx* stored in KZGSettings.tables (8th pointer slot)."""
        return C.c_void_p(int.from_bytes(self._s.raw[56:64], "little"))

    def launch_count(self):
        return int(lib().ckzg_b200_launch_count(self.engine))

    def close(self):
        if self._s is not None:
            lib().free_trusted_setup(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_trusted_setup(path=SETUP_TXT, precompute=0):
    return TrustedSetup(path, precompute)


def _need(b, n, what):
    b = bytes(b)
    if len(b) != n:
        raise ValueError("%s must be %d bytes" % (what, n))
    return b


def blob_to_kzg_commitment(blob, ts):
    out = C.create_string_buffer(48)
    _raise(lib().blob_to_kzg_commitment(out, _need(blob, BYTES_PER_BLOB, "blob"), ts.ptr), "blob_to_kzg_commitment")
    return out.raw


def compute_kzg_proof(blob, z, ts):
    proof, y = C.create_string_buffer(48), C.create_string_buffer(32)
    _raise(lib().compute_kzg_proof(proof, y, _need(blob, BYTES_PER_BLOB, "blob"), _need(z, 32, "z"), ts.ptr), "compute_kzg_proof")
    return proof.raw, y.raw


def compute_blob_kzg_proof(blob, commitment, ts):
    out = C.create_string_buffer(48)
    _raise(
        lib().compute_blob_kzg_proof(out, _need(blob, BYTES_PER_BLOB, "blob"), _need(commitment, 48, "commitment"), ts.ptr),
        "compute_blob_kzg_proof",
    )
    return out.raw


def verify_kzg_proof(commitment, z, y, proof, ts):
    ok = C.c_bool(False)
    _raise(
        lib().verify_kzg_proof(C.byref(ok), _need(commitment, 48, "commitment"), _need(z, 32, "z"), _need(y, 32, "y"), _need(proof, 48, "proof"), ts.ptr),
        "verify_kzg_proof",
    )
    return bool(ok.value)


def verify_blob_kzg_proof(blob, commitment, proof, ts):
    ok = C.c_bool(False)
    _raise(
        lib().verify_blob_kzg_proof(C.byref(ok), _need(blob, BYTES_PER_BLOB, "blob"), _need(commitment, 48, "commitment"), _need(proof, 48, "proof"), ts.ptr),
        "verify_blob_kzg_proof",
    )
    return bool(ok.value)


def verify_blob_kzg_proof_batch(blobs, commitments, proofs, ts):
    blobs, commitments, proofs = bytes(blobs), bytes(commitments), bytes(proofs)
    if len(blobs) % BYTES_PER_BLOB or len(commitments) % 48 or len(proofs) % 48:
        raise ValueError("inputs must be multiples of their element size")
    n = len(blobs) // BYTES_PER_BLOB
    if len(commitments) // 48 != n or len(proofs) // 48 != n:
        raise ValueError("expected same number of blobs/commitments/proofs")
    ok = C.c_bool(False)
    _raise(lib().verify_blob_kzg_proof_batch(C.byref(ok), blobs, commitments, proofs, C.c_uint64(n), ts.ptr), "verify_blob_kzg_proof_batch")
    return bool(ok.value)


# ---- engine-level batched calls (include/ckzg_b200.h) ---------------------------------------------


def blob_to_kzg_commitment_batch(blobs, ts, status=False):
    """n concatenated blobs (host bytes) -> n*48 bytes; one engine call."""
    blobs = bytes(blobs)
    n = len(blobs) // BYTES_PER_BLOB
    out = C.create_string_buffer(48 * max(n, 1))
    st = (C.c_int * max(n, 1))()
    rc = lib().ckzg_b200_blob_to_kzg_commitment_batch(ts.engine, out, blobs, C.c_uint64(n), HOST, st)
    if status:
        return out.raw[: 48 * n], list(st)[:n]
    _raise(rc, "ckzg_b200_blob_to_kzg_commitment_batch")
    return out.raw[: 48 * n]


def blob_to_kzg_commitment_device(out_ptr, blobs_ptr, n, ts):
    """Device-resident variant: raw CUDA pointers (e.g. torch tensor .data_ptr())."""
    _raise(
        lib().ckzg_b200_blob_to_kzg_commitment_batch(ts.engine, C.c_void_p(out_ptr), C.c_void_p(blobs_ptr), C.c_uint64(n), DEVICE, None),
        "ckzg_b200_blob_to_kzg_commitment_batch",
    )


def blob_to_kzg_commitment_batch_host(out_ptr, blobs_ptr, n, ts):
    """Same engine call with HOST pointers given as integers (pinned torch tensors, numpy arrays)."""
    _raise(
        lib().ckzg_b200_blob_to_kzg_commitment_batch(ts.engine, C.c_void_p(out_ptr), C.c_void_p(blobs_ptr), C.c_uint64(n), HOST, None),
        "ckzg_b200_blob_to_kzg_commitment_batch",
    )


def mulbench(ilp, iters, blocks, threads):
    """Fp Montgomery-multiplier throughput probe (include/ckzg_b200.h): ms for blocks x threads x ilp x iters products."""
    ms = C.c_float(0)
    _raise(lib().ckzg_b200_selftest_mulbench(C.c_int(ilp), C.c_int(iters), C.c_int(blocks), C.c_int(threads), C.byref(ms)), "ckzg_b200_selftest_mulbench")
    return float(ms.value)


def compute_blob_kzg_proof_device(out_ptr, blobs_ptr, commitments_ptr, n, ts):
    _raise(
        lib().ckzg_b200_compute_blob_kzg_proof_batch(ts.engine, C.c_void_p(out_ptr), C.c_void_p(blobs_ptr), C.c_void_p(commitments_ptr), C.c_uint64(n), DEVICE, None),
        "ckzg_b200_compute_blob_kzg_proof_batch",
    )


def verify_blob_kzg_proof_batch_host(blobs_ptr, commitments_ptr, proofs_ptr, n, ts):
    """Same engine call with HOST pointers given as integers (e.g. pinned torch tensors)."""
    ok = C.c_int(0)
    _raise(
        lib().ckzg_b200_verify_blob_kzg_proof_batch(ts.engine, C.byref(ok), C.c_void_p(blobs_ptr), C.c_void_p(commitments_ptr), C.c_void_p(proofs_ptr), C.c_uint64(n), HOST),
        "ckzg_b200_verify_blob_kzg_proof_batch",
    )
    return bool(ok.value)


class VerifyShard:
    """One rank's shard of a sharded verify_blob_kzg_proof_batch (include/ckzg_b200.h, parallel.py): the per-blob
    stage runs in the constructor (-> self.zy, n x 64 bytes of z || y), the validated points, their vmsm table, z and y
    stay on the device until stage2()."""

    def __init__(self, blobs_ptr, commitments_ptr, proofs_ptr, n, ts, mem=DEVICE):
        self.n = n
        self._h = C.c_void_p(None)
        out = C.create_string_buffer(64 * max(n, 1))
        _raise(
            lib().ckzg_b200_verify_shard_stage1(ts.engine, C.byref(self._h), out, C.c_void_p(blobs_ptr), C.c_void_p(commitments_ptr), C.c_void_p(proofs_ptr), C.c_uint64(n), mem),
            "ckzg_b200_verify_shard_stage1",
        )
        self.zy = out.raw[: 64 * n]

    def stage2(self, tuples, n_total, first):
        part = C.create_string_buffer(384)
        _raise(lib().ckzg_b200_verify_shard_stage2(self._h, part, bytes(tuples), C.c_uint64(n_total), C.c_uint64(first)), "ckzg_b200_verify_shard_stage2")
        return part.raw

    def close(self):
        if self._h:
            lib().ckzg_b200_verify_shard_free(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_verify_tuples(commitments, zy, proofs, n):
    """n x 160-byte records C || z || y || proof (eip4844.c:648-660), assembled in C."""
    out = C.create_string_buffer(160 * max(n, 1))
    _raise(lib().ckzg_b200_pack_verify_tuples(out, bytes(commitments), bytes(zy), bytes(proofs), C.c_uint64(n)), "ckzg_b200_pack_verify_tuples")
    return out.raw[: 160 * n]


def verify_shard_finish(partials, n_ranks, ts):
    ok = C.c_int(0)
    _raise(lib().ckzg_b200_verify_shard_finish(ts.engine, C.byref(ok), bytes(partials), C.c_uint64(n_ranks)), "ckzg_b200_verify_shard_finish")
    return bool(ok.value)


def set_caller_stream(cuda_stream_ptr):
    """DEVICE-mode ordering (include/ckzg_b200.h): this thread's engine calls wait for work enqueued on this stream."""
    lib().ckzg_b200_set_caller_stream(C.c_void_p(cuda_stream_ptr or None))


def compute_cells_and_kzg_proofs(blob, ts):
    """-> (cells[128], proofs[128]) as in bindings/python/ckzg_wrap.c"""
    cells = C.create_string_buffer(CELLS_PER_EXT_BLOB * BYTES_PER_CELL)
    proofs = C.create_string_buffer(CELLS_PER_EXT_BLOB * 48)
    lib().compute_cells_and_kzg_proofs.restype = C.c_int
    _raise(lib().compute_cells_and_kzg_proofs(cells, proofs, _need(blob, BYTES_PER_BLOB, "blob"), ts.ptr), "compute_cells_and_kzg_proofs")
    return ([cells.raw[BYTES_PER_CELL * k : BYTES_PER_CELL * (k + 1)] for k in range(128)], [proofs.raw[48 * k : 48 * k + 48] for k in range(128)])


def recover_cells_and_kzg_proofs(cell_indices, cells, ts):
    n = len(cell_indices)
    if len(cells) != n:
        raise ValueError("expected same number of indices and cells")
    idx = (C.c_uint64 * max(n, 1))(*cell_indices)
    out_c = C.create_string_buffer(CELLS_PER_EXT_BLOB * BYTES_PER_CELL)
    out_p = C.create_string_buffer(CELLS_PER_EXT_BLOB * 48)
    lib().recover_cells_and_kzg_proofs.restype = C.c_int
    _raise(
        lib().recover_cells_and_kzg_proofs(out_c, out_p, idx, b"".join(_need(c, BYTES_PER_CELL, "cell") for c in cells), C.c_uint64(n), ts.ptr),
        "recover_cells_and_kzg_proofs",
    )
    return ([out_c.raw[BYTES_PER_CELL * k : BYTES_PER_CELL * (k + 1)] for k in range(128)], [out_p.raw[48 * k : 48 * k + 48] for k in range(128)])


def verify_cell_kzg_proof_batch(commitments, cell_indices, cells, proofs, ts):
    n = len(cell_indices)
    if not (len(commitments) == len(cells) == len(proofs) == n):
        raise ValueError("expected same number of commitments, indices, cells and proofs")
    idx = (C.c_uint64 * max(n, 1))(*cell_indices)
    ok = C.c_bool(False)
    lib().verify_cell_kzg_proof_batch.restype = C.c_int
    _raise(
        lib().verify_cell_kzg_proof_batch(
            C.byref(ok), b"".join(_need(c, 48, "commitment") for c in commitments), idx, b"".join(_need(c, BYTES_PER_CELL, "cell") for c in cells),
            b"".join(_need(p, 48, "proof") for p in proofs), C.c_uint64(n), ts.ptr,
        ),
        "verify_cell_kzg_proof_batch",
    )
    return bool(ok.value)


def compute_cells_and_kzg_proofs_device(cells_ptr, proofs_ptr, blobs_ptr, n, ts):
    """Batched, device pointers (0 = skip that output)."""
    lib().ckzg_b200_compute_cells_and_kzg_proofs_batch.restype = C.c_int
    _raise(
        lib().ckzg_b200_compute_cells_and_kzg_proofs_batch(ts.engine, C.c_void_p(cells_ptr or None), C.c_void_p(proofs_ptr or None), C.c_void_p(blobs_ptr), C.c_uint64(n), DEVICE, None),
        "ckzg_b200_compute_cells_and_kzg_proofs_batch",
    )


def compute_cells_and_kzg_proofs_host(cells_ptr, proofs_ptr, blobs_ptr, n, ts):
    lib().ckzg_b200_compute_cells_and_kzg_proofs_batch.restype = C.c_int
    _raise(
        lib().ckzg_b200_compute_cells_and_kzg_proofs_batch(ts.engine, C.c_void_p(cells_ptr or None), C.c_void_p(proofs_ptr or None), C.c_void_p(blobs_ptr), C.c_uint64(n), HOST, None),
        "ckzg_b200_compute_cells_and_kzg_proofs_batch",
    )


def recover_cells_and_kzg_proofs_device(cells_out_ptr, proofs_out_ptr, cell_indices, cells_ptr, num_cells, n, ts):
    idx = (C.c_uint64 * (n * num_cells))(*cell_indices)
    lib().ckzg_b200_recover_cells_and_kzg_proofs_batch.restype = C.c_int
    _raise(
        lib().ckzg_b200_recover_cells_and_kzg_proofs_batch(ts.engine, C.c_void_p(cells_out_ptr), C.c_void_p(proofs_out_ptr or None), idx, C.c_void_p(cells_ptr), C.c_uint64(num_cells), C.c_uint64(n), DEVICE, None),
        "ckzg_b200_recover_cells_and_kzg_proofs_batch",
    )


def profile_enable(ts, level=2):
    """0 off, 1 whole-call device time only, 2 per-kernel times (stages serialised)."""
    lib().ckzg_b200_profile_enable.restype = None
    lib().ckzg_b200_profile_enable(ts.engine, C.c_int(int(level)))


def profile_dump(ts):
    import json

    buf = C.create_string_buffer(1 << 16)
    lib().ckzg_b200_profile_dump.restype = C.c_int
    lib().ckzg_b200_profile_dump(ts.engine, buf, C.c_size_t(len(buf)))
    return json.loads(buf.value.decode())


def verify_blob_kzg_proof_batch_device(blobs_ptr, commitments_ptr, proofs_ptr, n, ts):
    ok = C.c_int(0)
    _raise(
        lib().ckzg_b200_verify_blob_kzg_proof_batch(ts.engine, C.byref(ok), C.c_void_p(blobs_ptr), C.c_void_p(commitments_ptr), C.c_void_p(proofs_ptr), C.c_uint64(n), DEVICE),
        "ckzg_b200_verify_blob_kzg_proof_batch",
    )
    return bool(ok.value)


def verify_cell_kzg_proof_batch_ptr(commitments_ptr, cell_indices, cells_ptr, proofs_ptr, n, ts, mem=HOST):
    """Raw-pointer form of verify_cell_kzg_proof_batch (host or device buffers; `cell_indices` is a ctypes
    uint64 array or any sequence, always host memory as in the C ABI)."""
    idx = cell_indices if isinstance(cell_indices, C.Array) else (C.c_uint64 * max(n, 1))(*cell_indices)
    ok = C.c_int(0)
    lib().ckzg_b200_verify_cell_kzg_proof_batch.restype = C.c_int
    _raise(
        lib().ckzg_b200_verify_cell_kzg_proof_batch(ts.engine, C.byref(ok), C.c_void_p(commitments_ptr), idx, C.c_void_p(cells_ptr), C.c_void_p(proofs_ptr), C.c_uint64(n), mem),
        "ckzg_b200_verify_cell_kzg_proof_batch",
    )
    return bool(ok.value)


def coalesce_stats(ts):
    """{op: (requests, batches, largest batch)} of the call-coalescing front end (ckzg_b200_coalesce_stats)."""
    out = (C.c_uint64 * 15)()
    _raise(lib().ckzg_b200_coalesce_stats(ts.engine, out), "ckzg_b200_coalesce_stats")
    names = ("blob_to_kzg_commitment", "compute_blob_kzg_proof", "compute_kzg_proof", "compute_cells_and_kzg_proofs", "recover_cells_and_kzg_proofs")
    return {n: (int(out[3 * i]), int(out[3 * i + 1]), int(out[3 * i + 2])) for i, n in enumerate(names)}


def coalesce_enable(ts, on=True):
    _raise(lib().ckzg_b200_coalesce_enable(ts.engine, C.c_int(1 if on else 0)), "ckzg_b200_coalesce_enable")


def bench_per_blob_callers(ts, op, threads, reps, blobs_ptr, n_blobs):
    """blobs/s of `threads` native host threads calling the per-blob API (0 = commitment, 1 = cells + proofs)."""
    sec = C.c_double(0)
    _raise(lib().ckzg_b200_bench_per_blob_callers(ts.engine, C.c_int(op), C.c_int(threads), C.c_int(reps), C.c_void_p(blobs_ptr), C.c_uint64(n_blobs), C.byref(sec)), "ckzg_b200_bench_per_blob_callers")
    return threads * reps / sec.value
