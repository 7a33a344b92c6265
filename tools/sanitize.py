#!/usr/bin/env python3
"""Small end-to-end pass over every API entry for compute-sanitizer (run under gpurun):
    compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

mod = entry.load_package()
ts = mod.load_trusted_setup()
blobs = bench.synth_blobs(3, 11).tobytes()
bl = [blobs[131072 * i : 131072 * (i + 1)] for i in range(3)]
cms = [mod.blob_to_kzg_commitment(b, ts) for b in bl]
prs = [mod.compute_blob_kzg_proof(b, c, ts) for b, c in zip(bl, cms)]
p, y = mod.compute_kzg_proof(bl[0], (12345).to_bytes(32, "big"), ts)
assert mod.verify_kzg_proof(cms[0], (12345).to_bytes(32, "big"), y, p, ts)
assert mod.verify_blob_kzg_proof(bl[0], cms[0], prs[0], ts)
assert mod.verify_blob_kzg_proof_batch(b"".join(bl), b"".join(cms), b"".join(prs), ts)
cells, proofs = mod.compute_cells_and_kzg_proofs(bl[1], ts)
idx = list(range(0, 128, 2))
rc, rp = mod.recover_cells_and_kzg_proofs(idx, [cells[i] for i in idx], ts)
assert rc == cells and rp == proofs
assert mod.verify_cell_kzg_proof_batch([cms[1]] * 128, list(range(128)), cells, proofs, ts)
print("sanitize pass ok")
