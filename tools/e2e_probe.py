import os, sys, json, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import __graft_entry__ as entry, bench
mod = entry.load_package(); ts = mod.load_trusted_setup()
n = int(os.environ.get("PROBE_N", "4096"))
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
print(os.environ.get('CKZG_B200_TAIL_PIECES'), 'wall ms', round(dt*1e3, 3), {k: round(v[0]/max(1,v[1]), 3) for k, v in p['kernels'].items() if k.startswith('stage') or k.startswith('host')}, 'engine', round(p['call_ms']/p['calls'], 3))
if os.environ.get("PROBE_TIMERS"):
    import ctypes as C
    buf = torch.zeros(2 + 2 * 4096, dtype=torch.int32, device="cuda")
    buf[1] = 4096
    lib = mod.lib()
    assert lib.ckzg_b200_debug_timers(C.c_void_p(buf.data_ptr())) == 0
    torch.cuda.synchronize()
    mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)
    torch.cuda.synchronize()
    lib.ckzg_b200_debug_timers(C.c_void_p(0))
    b = buf.cpu().numpy().astype("uint32")
    recs = [(int(b[2 + 2 * i]), int(b[3 + 2 * i])) for i in range(int(b[0]))]
    t0 = min(t for _, t in recs)
    names = {1: "hash", 2: "validate", 3: "evaluate"}
    out = sorted(((t - t0) & 0xffffffff, i) for i, t in recs)
    print("device timeline (us since the first record; hash:<first block>):", ", ".join("%s%s%s@%d" % (names.get(i & 0x7f, "?"), ":%d" % (i >> 8) if (i & 0x7f) == 1 else "", " end" if i & 0x80 else "", dt // 1000) for dt, i in out))
if os.environ.get("PROBE_CALLERS"):
    import threading
    k = int(os.environ["PROBE_CALLERS"])
    def worker():
        for _ in range(6):
            mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)
    th = [threading.Thread(target=worker) for _ in range(k)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    print("%d concurrent callers: %.0f blobs/s end to end" % (k, k * 6 * n / (time.perf_counter() - t0)))
