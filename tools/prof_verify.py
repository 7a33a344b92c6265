#!/usr/bin/env python3
"""Driver for ncu captures of the blob-verification kernels: three device-resident
verify_blob_kzg_proof_batch calls of N blobs (default 4096).  Usage under gpurun:
  ncu --set full --import-source on -k regex:<kernel> -s 1 -c 1 -o gpurun_out/prof python tools/prof_verify.py [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mod = entry.load_package()
ts = mod.load_trusted_setup()
blobs = torch.from_numpy(bench.synth_blobs(n, 9)).cuda()
cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
mod.blob_to_kzg_commitment_device(cms.data_ptr(), blobs.data_ptr(), n, ts)
mod.compute_blob_kzg_proof_device(prs.data_ptr(), blobs.data_ptr(), cms.data_ptr(), n, ts)
for _ in range(3):
    assert mod.verify_blob_kzg_proof_batch_device(blobs.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)
torch.cuda.synchronize()
print("ok")
