#!/usr/bin/env python3
"""blob_to_kzg_commitment batch N (device-resident): per-kernel times and blobs/s, and a checksum of the
commitments (compare across CKZG_B200_COMMIT_WINDOW settings: 0 = bucket MSM).  Usage under gpurun:
  python tools/time_commit.py [N]"""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mod = entry.load_package()
t0 = time.perf_counter()
ts = mod.load_trusted_setup()
print("load_trusted_setup %.2f s  (CKZG_B200_COMMIT_WINDOW=%s)" % (time.perf_counter() - t0, os.environ.get("CKZG_B200_COMMIT_WINDOW", "auto")))
blobs = torch.from_numpy(bench.synth_blobs(n, 9)).cuda()
out = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
mod.blob_to_kzg_commitment_device(out.data_ptr(), blobs.data_ptr(), n, ts)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    mod.blob_to_kzg_commitment_device(out.data_ptr(), blobs.data_ptr(), n, ts)
dt = (time.perf_counter() - t0) / 3
print("batch %d: %.2f ms per call, %.0f blobs/s" % (n, dt * 1e3, n / dt))
mod.profile_enable(ts, 2)
for _ in range(3):
    mod.blob_to_kzg_commitment_device(out.data_ptr(), blobs.data_ptr(), n, ts)
p = mod.profile_dump(ts)
mod.profile_enable(ts, 0)
print({k: round(v[0] / 3, 3) for k, v in p["kernels"].items() if k not in ("begin", "end")})
print("commitments sha256", hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16])
one = bytes(blobs[:131072].cpu().numpy().tobytes())
mod.blob_to_kzg_commitment(one, ts)
t0 = time.perf_counter()
for _ in range(5):
    mod.blob_to_kzg_commitment(one, ts)
print("single call %.3f ms" % (1000 * (time.perf_counter() - t0) / 5))
