#!/usr/bin/env python3
"""Driver for ncu captures of the EIP-7594 kernels: two compute_cells_and_kzg_proofs batches of
N blobs (default 256), device-resident.  Usage under gpurun:
  ncu --set full --import-source on -k regex:g1_fft_stage -s 20 -c 1 -o gpurun_out/prof python tools/prof_cells.py [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mod = entry.load_package()
ts = mod.load_trusted_setup()
blobs = torch.from_numpy(bench.synth_blobs(n, 9)).cuda()
cells = torch.empty(n * 128 * 2048, dtype=torch.uint8, device="cuda")
proofs = torch.empty(n * 128 * 48, dtype=torch.uint8, device="cuda")
for _ in range(2):
    mod.compute_cells_and_kzg_proofs_device(cells.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, ts)
torch.cuda.synchronize()
print("ok")
