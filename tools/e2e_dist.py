#!/usr/bin/env python3
"""Per-call wall times of the end-to-end (pinned host pointers) verify_blob_kzg_proof_batch call, n = 4096: is the
distribution tight, or do some calls fall into a slow mode?  Usage under gpurun: python tools/e2e_dist.py [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n = int(os.environ.get("PROBE_N", "4096"))
mod = entry.load_package()
ts = mod.load_trusted_setup()
host = torch.from_numpy(bench.synth_blobs(n, 1)).pin_memory()
dev = host.cuda()
cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
hc, hp = cms.cpu().pin_memory(), prs.cpu().pin_memory()
for _ in range(5):
    assert mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)
t = []
for _ in range(calls):
    t0 = time.perf_counter()
    mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)
    t.append(1e3 * (time.perf_counter() - t0))
s = sorted(t)
print("e2e per-call ms over %d calls: min %.2f  p25 %.2f  median %.2f  p75 %.2f  max %.2f  mean %.2f" % (calls, s[0], s[len(s) // 4], s[len(s) // 2], s[3 * len(s) // 4], s[-1], sum(t) / len(t)))
print("all:", " ".join("%.1f" % x for x in t))
# the same call with the device-resident inputs, for the host overhead around the engine
t = []
for _ in range(20):
    t0 = time.perf_counter()
    mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)
    t.append(1e3 * (time.perf_counter() - t0))
print("device-resident per-call ms: min %.2f median %.2f max %.2f" % (min(t), sorted(t)[10], max(t)))
