"""ctypes front end shared by the reference build (oracle/_ref/libckzg_ref.so) and by any other
library that exports the c-kzg-4844 C API (the product .so exports the same symbols, so tests can
drive both through this one wrapper).

ORACLE / TEST INFRASTRUCTURE ONLY when it points at oracle/_ref.

API mirrored: src/eip4844/eip4844.h:43-81, src/eip7594/eip7594.h:35-56, src/setup/setup.h:31-44.
Return codes: src/common/ret.h:24-29.
"""
import ctypes as C
import os

C_KZG_OK, C_KZG_BADARGS, C_KZG_ERROR, C_KZG_MALLOC = 0, 1, 2, 3
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libckzg_ref.so")
SETUP_TXT = os.path.join(os.path.dirname(_HERE), "c-kzg-4844_b200", "data", "trusted_setup.txt")


class KzgError(Exception):
    def __init__(self, code, fn):
        super().__init__("%s -> C_KZG_RET %d" % (fn, code))
        self.code = code


class BadArgs(KzgError):
    pass


def _check(code, fn):
    if code == C_KZG_BADARGS:
        raise BadArgs(code, fn)
    if code != C_KZG_OK:
        raise KzgError(code, fn)


def parse_trusted_setup_text(path):
    """-> (g1_monomial_bytes, g1_lagrange_bytes, g2_monomial_bytes); format src/setup/setup.c:516-582"""
    with open(path) as f:
        tok = f.read().split()
    n1, n2 = int(tok[0]), int(tok[1])
    body = tok[2:]
    lag = bytes.fromhex("".join(body[:n1]))
    g2 = bytes.fromhex("".join(body[n1 : n1 + n2]))
    mono = bytes.fromhex("".join(body[n1 + n2 : n1 + n2 + n1]))
    return mono, lag, g2


class CKZG:
    """One loaded library + one trusted setup."""

    def __init__(self, so_path=REF_SO, setup_path=SETUP_TXT, precompute=0):
        if not os.path.exists(so_path):
            raise FileNotFoundError(so_path)
        self.lib = C.CDLL(so_path)
        self.settings = C.create_string_buffer(80)  # sizeof(KZGSettings), src/setup/settings.h:27-79
        mono, lag, g2 = parse_trusted_setup_text(setup_path)
        fn = self.lib.load_trusted_setup
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_uint64]
        _check(fn(self.settings, mono, len(mono), lag, len(lag), g2, len(g2), precompute), "load_trusted_setup")
        self.precompute = precompute
        for name in (
            "blob_to_kzg_commitment compute_kzg_proof compute_blob_kzg_proof verify_kzg_proof "
            "verify_blob_kzg_proof verify_blob_kzg_proof_batch compute_cells_and_kzg_proofs "
            "recover_cells_and_kzg_proofs verify_cell_kzg_proof_batch"
        ).split():
            getattr(self.lib, name).restype = C.c_int

    def close(self):
        if self.settings is not None:
            self.lib.free_trusted_setup.restype = None
            self.lib.free_trusted_setup(self.settings)
            self.settings = None

    # --- EIP-4844 ---------------------------------------------------------------------------
    def blob_to_kzg_commitment(self, blob):
        out = C.create_string_buffer(48)
        _check(self.lib.blob_to_kzg_commitment(out, bytes(blob), self.settings), "blob_to_kzg_commitment")
        return out.raw

    def compute_kzg_proof(self, blob, z):
        proof, y = C.create_string_buffer(48), C.create_string_buffer(32)
        _check(self.lib.compute_kzg_proof(proof, y, bytes(blob), bytes(z), self.settings), "compute_kzg_proof")
        return proof.raw, y.raw

    def compute_blob_kzg_proof(self, blob, commitment):
        out = C.create_string_buffer(48)
        _check(self.lib.compute_blob_kzg_proof(out, bytes(blob), bytes(commitment), self.settings), "compute_blob_kzg_proof")
        return out.raw

    def verify_kzg_proof(self, commitment, z, y, proof):
        ok = C.c_bool(False)
        _check(self.lib.verify_kzg_proof(C.byref(ok), bytes(commitment), bytes(z), bytes(y), bytes(proof), self.settings), "verify_kzg_proof")
        return bool(ok.value)

    def verify_blob_kzg_proof(self, blob, commitment, proof):
        ok = C.c_bool(False)
        _check(self.lib.verify_blob_kzg_proof(C.byref(ok), bytes(blob), bytes(commitment), bytes(proof), self.settings), "verify_blob_kzg_proof")
        return bool(ok.value)

    def verify_blob_kzg_proof_batch(self, blobs, commitments, proofs):
        """blobs/commitments/proofs: concatenated bytes (n inferred from commitments)."""
        n = len(commitments) // 48
        ok = C.c_bool(False)
        _check(
            self.lib.verify_blob_kzg_proof_batch(C.byref(ok), bytes(blobs), bytes(commitments), bytes(proofs), C.c_uint64(n), self.settings),
            "verify_blob_kzg_proof_batch",
        )
        return bool(ok.value)

    # --- EIP-7594 ---------------------------------------------------------------------------
    def compute_cells_and_kzg_proofs(self, blob, want_cells=True, want_proofs=True):
        cells = C.create_string_buffer(128 * 2048) if want_cells else None
        proofs = C.create_string_buffer(128 * 48) if want_proofs else None
        _check(self.lib.compute_cells_and_kzg_proofs(cells, proofs, bytes(blob), self.settings), "compute_cells_and_kzg_proofs")
        return (cells.raw if want_cells else None), (proofs.raw if want_proofs else None)

    def recover_cells_and_kzg_proofs(self, cell_indices, cells, want_proofs=True):
        n = len(cell_indices)
        idx = (C.c_uint64 * n)(*cell_indices)
        out_cells = C.create_string_buffer(128 * 2048)
        out_proofs = C.create_string_buffer(128 * 48) if want_proofs else None
        _check(
            self.lib.recover_cells_and_kzg_proofs(out_cells, out_proofs, idx, bytes(cells), C.c_uint64(n), self.settings),
            "recover_cells_and_kzg_proofs",
        )
        return out_cells.raw, (out_proofs.raw if want_proofs else None)

    def verify_cell_kzg_proof_batch(self, commitments, cell_indices, cells, proofs):
        n = len(cell_indices)
        idx = (C.c_uint64 * n)(*cell_indices)
        ok = C.c_bool(False)
        _check(
            self.lib.verify_cell_kzg_proof_batch(C.byref(ok), bytes(commitments), idx, bytes(cells), bytes(proofs), C.c_uint64(n), self.settings),
            "verify_cell_kzg_proof_batch",
        )
        return bool(ok.value)
