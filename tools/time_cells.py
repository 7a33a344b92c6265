#!/usr/bin/env python3
"""Per-kernel times of compute_cells_and_kzg_proofs (batch N, device-resident, 3 calls) checked against a
reference batch computed once.  Usage under gpurun: python tools/time_cells.py [N]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mod = entry.load_package()
t0 = time.perf_counter()
ts = mod.load_trusted_setup()
print("load_trusted_setup %.2f s" % (time.perf_counter() - t0))
blobs = torch.from_numpy(bench.synth_blobs(n, 9)).cuda()
cells = torch.empty(n * 128 * 2048, dtype=torch.uint8, device="cuda")
proofs = torch.empty(n * 128 * 48, dtype=torch.uint8, device="cuda")
mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, ts)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, ts)
dt = (time.perf_counter() - t0) / 3
print("batch %d: %.2f ms per call, %.0f blobs/s" % (n, dt * 1e3, n / dt))
mod.profile_enable(ts, 2)
for _ in range(3):
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, ts)
p = mod.profile_dump(ts)
mod.profile_enable(ts, 0)
print({k: round(v[0] / 3, 3) for k, v in p["kernels"].items() if k not in ("begin", "end")})
import hashlib

print("proofs sha256", hashlib.sha256(proofs.cpu().numpy().tobytes()).hexdigest()[:16])
