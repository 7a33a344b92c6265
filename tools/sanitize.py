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

# ---- round 2: batched / device / pinned / multi-device paths at sizes the sanitizer can afford ----
import ctypes as C  # noqa: E402

import torch  # noqa: E402

n = int(os.environ.get("SANITIZE_N", "1024"))
host = torch.from_numpy(bench.synth_blobs(n, 12))
dev = host.cuda()
d_cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
d_prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
mod.blob_to_kzg_commitment_device(d_cms.data_ptr(), dev.data_ptr(), n, ts)
mod.compute_blob_kzg_proof_device(d_prs.data_ptr(), dev.data_ptr(), d_cms.data_ptr(), n, ts)
# device-resident batch: fused stage 1, chunked evaluations with the streamed transcript, bucket MSMs, two-machine pairing
assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), d_cms.data_ptr(), d_prs.data_ptr(), n, ts)
# pinned host batch: chunked upload, (with CKZG_B200_TAIL_PIECES=8) the tail in column pieces with resumed hashes
pin, hc, hp = host.pin_memory(), d_cms.cpu().pin_memory(), d_prs.cpu().pin_memory()
assert mod.verify_blob_kzg_proof_batch_host(pin.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)
# pageable host batch: staged through the pinned ring by the host threads
pg = host.numpy().copy()
assert mod.verify_blob_kzg_proof_batch_host(pg.ctypes.data, hc.data_ptr(), hp.data_ptr(), n, ts)
# FK20 with one thread per butterfly (forced on a small batch) against the quad form
m = 8
c1 = torch.empty(m * 262144, dtype=torch.uint8, device="cuda")
p1 = torch.empty(m * 6144, dtype=torch.uint8, device="cuda")
c2, p2 = torch.empty_like(c1), torch.empty_like(p1)
mod.compute_cells_and_kzg_proofs_device(c1.data_ptr(), p1.data_ptr(), dev.data_ptr(), m, ts)
os.environ["CKZG_B200_FFT_THREAD_MIN"] = "1"
mod.compute_cells_and_kzg_proofs_device(c2.data_ptr(), p2.data_ptr(), dev.data_ptr(), m, ts)
del os.environ["CKZG_B200_FFT_THREAD_MIN"]
assert torch.equal(c1, c2) and torch.equal(p1, p2)
# one context over two replicas of device 0: sharded verification with one challenge, sub-batched cell verification
os.environ["CKZG_B200_DEVICES"] = "0,0"
ts2 = mod.load_trusted_setup()
del os.environ["CKZG_B200_DEVICES"]
k = min(n, 512)
assert mod.verify_blob_kzg_proof_batch_host(pg.ctypes.data, hc.data_ptr(), hp.data_ptr(), k, ts2)
ts2.close()
print("sanitize pass ok (round-2 paths, n = %d)" % n)
