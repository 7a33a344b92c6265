import os, sys, json, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import __graft_entry__ as entry, bench
mod = entry.load_package(); ts = mod.load_trusted_setup()
n = 4096
host = torch.from_numpy(bench.synth_blobs(n, 1)).pin_memory(); dev = host.cuda()
cms = torch.empty(48*n, dtype=torch.uint8, device='cuda'); prs = torch.empty(48*n, dtype=torch.uint8, device='cuda')
mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts); mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
hc, hp = cms.cpu().pin_memory(), prs.cpu().pin_memory()
for _ in range(4): assert mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)
mod.profile_enable(ts, 1)
t0 = time.perf_counter()
for _ in range(10): mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)
dt = (time.perf_counter() - t0) / 10
p = mod.profile_dump(ts); mod.profile_enable(ts, 0)
print(os.environ.get('CKZG_B200_TAIL_PIECES'), 'wall ms', round(dt*1e3, 3), {k: round(v[0]/max(1,v[1]), 3) for k, v in p['kernels'].items() if k.startswith('stage')}, 'engine', round(p['call_ms']/p['calls'], 3))
