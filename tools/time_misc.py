#!/usr/bin/env python3
"""Per-kernel times of compute_blob_kzg_proof (batch 1024) and recover_cells_and_kzg_proofs (batch 256, every
other cell missing), device-resident.  Usage under gpurun: python tools/time_misc.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

mod = entry.load_package()
ts = mod.load_trusted_setup()
n = 1024
blobs = torch.from_numpy(bench.synth_blobs(n, 9)).cuda()
cms = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
prs = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
mod.blob_to_kzg_commitment_device(cms.data_ptr(), blobs.data_ptr(), n, ts)
mod.compute_blob_kzg_proof_device(prs.data_ptr(), blobs.data_ptr(), cms.data_ptr(), n, ts)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    mod.compute_blob_kzg_proof_device(prs.data_ptr(), blobs.data_ptr(), cms.data_ptr(), n, ts)
dt = (time.perf_counter() - t0) / 3
print("compute_blob_kzg_proof batch %d: %.2f ms per call, %.0f blobs/s" % (n, dt * 1e3, n / dt))
mod.profile_enable(ts, 2)
for _ in range(3):
    mod.compute_blob_kzg_proof_device(prs.data_ptr(), blobs.data_ptr(), cms.data_ptr(), n, ts)
p = mod.profile_dump(ts)
mod.profile_enable(ts, 0)
print({k: round(v[0] / 3, 3) for k, v in p["kernels"].items() if k not in ("begin", "end")})
assert mod.verify_blob_kzg_proof_batch_device(blobs.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)

m = 256
cells = torch.empty(m * 262144, dtype=torch.uint8, device="cuda")
cprf = torch.empty(m * 128 * 48, dtype=torch.uint8, device="cuda")
mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), cprf.data_ptr(), blobs.data_ptr(), m, ts)
idx = list(range(0, 128, 2)) * m
given = cells.view(m, 128, 2048)[:, 0::2, :].contiguous()
rc = torch.empty_like(cells)
rp = torch.empty_like(cprf)
mod.recover_cells_and_kzg_proofs_device(rc.data_ptr(), rp.data_ptr(), idx, given.data_ptr(), 64, m, ts)
torch.cuda.synchronize()
assert torch.equal(rc, cells) and torch.equal(rp, cprf)
t0 = time.perf_counter()
for _ in range(3):
    mod.recover_cells_and_kzg_proofs_device(rc.data_ptr(), rp.data_ptr(), idx, given.data_ptr(), 64, m, ts)
dt = (time.perf_counter() - t0) / 3
print("recover batch %d: %.2f ms per call, %.0f blobs/s" % (m, dt * 1e3, m / dt))
mod.profile_enable(ts, 2)
for _ in range(3):
    mod.recover_cells_and_kzg_proofs_device(rc.data_ptr(), rp.data_ptr(), idx, given.data_ptr(), 64, m, ts)
p = mod.profile_dump(ts)
mod.profile_enable(ts, 0)
print({k: round(v[0] / 3, 3) for k, v in p["kernels"].items() if k not in ("begin", "end")})
t0 = time.perf_counter()
for _ in range(3):
    mod.recover_cells_and_kzg_proofs_device(rc.data_ptr(), 0, idx, given.data_ptr(), 64, m, ts)
dt = (time.perf_counter() - t0) / 3
print("recover cells only batch %d: %.2f ms per call, %.0f blobs/s" % (m, dt * 1e3, m / dt))
