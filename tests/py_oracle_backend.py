"""Adapter giving oracle/kzg_oracle.py (pure Python restatement) the CKZG method set."""
import os

from oracle import kzg_oracle as K

SETUP_TXT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "c-kzg-4844_b200", "data", "trusted_setup.txt")
BadArgs = K.BadArgs


class PyOracle:
    def __init__(self, setup_path=SETUP_TXT, check=False):
        self.s = K.load_trusted_setup_file(setup_path, check=check)

    def blob_to_kzg_commitment(self, blob):
        return K.blob_to_kzg_commitment(blob, self.s)

    def compute_kzg_proof(self, blob, z):
        return K.compute_kzg_proof(blob, z, self.s)

    def compute_blob_kzg_proof(self, blob, commitment):
        return K.compute_blob_kzg_proof(blob, commitment, self.s)

    def verify_kzg_proof(self, c, z, y, p):
        return K.verify_kzg_proof(c, z, y, p, self.s)

    def verify_blob_kzg_proof(self, blob, c, p):
        return K.verify_blob_kzg_proof(blob, c, p, self.s)

    def verify_blob_kzg_proof_batch(self, blobs, cs, ps):
        n = len(cs) // 48
        return K.verify_blob_kzg_proof_batch(
            [blobs[131072 * i : 131072 * (i + 1)] for i in range(n)],
            [cs[48 * i : 48 * i + 48] for i in range(n)],
            [ps[48 * i : 48 * i + 48] for i in range(n)],
            self.s,
        )

    def compute_cells_and_kzg_proofs(self, blob, want_cells=True, want_proofs=True):
        cells, proofs = K.compute_cells_and_kzg_proofs(blob, self.s, want_proofs=want_proofs)
        return cells, proofs

    def recover_cells_and_kzg_proofs(self, cell_indices, cells, want_proofs=True):
        n = len(cell_indices)
        return K.recover_cells_and_kzg_proofs(list(cell_indices), [cells[2048 * i : 2048 * (i + 1)] for i in range(n)], self.s, want_proofs=want_proofs)

    def verify_cell_kzg_proof_batch(self, commitments, cell_indices, cells, proofs):
        n = len(cell_indices)
        return K.verify_cell_kzg_proof_batch(
            [commitments[48 * i : 48 * i + 48] for i in range(n)], list(cell_indices), [cells[2048 * i : 2048 * (i + 1)] for i in range(n)],
            [proofs[48 * i : 48 * i + 48] for i in range(n)], self.s,
        )
