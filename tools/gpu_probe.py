#!/usr/bin/env python3
"""GPU-side probes (run under gpurun): Montgomery-multiplier throughput (the measured integer-pipe
roofline denominator) and MSM accumulate register-cap variants.  Prints JSON lines."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def mulbench():
    import __graft_entry__ as entry

    mod = entry.load_package()
    lib = mod.lib()
    lib.ckzg_b200_selftest_mulbench.restype = C.c_int
    out = []
    iters = 2000
    for threads in (128, 256):
        for bps in (1, 2, 4, 8):  # blocks per SM
            for ilp in (1, 2, 4):
                ms = C.c_float(0)
                blocks = 148 * bps
                rc = lib.ckzg_b200_selftest_mulbench(ilp, iters, blocks, threads, C.byref(ms))
                muls = blocks * threads * iters * ilp
                out.append({"threads": threads, "blocks_per_sm": bps, "ilp": ilp, "ms": ms.value, "fp_mul_per_s": muls / (ms.value * 1e-3),
                            "mac_per_s": 300 * muls / (ms.value * 1e-3), "rc": rc})
    best = max(out, key=lambda r: r["mac_per_s"])
    print(json.dumps({"probe": "mulbench", "best": best, "all": out}))


def lanes_and_dfma():
    """(1) One warp per SM sub-partition with only the first k lanes active: does a partial warp issue
    the wide multiplies faster (latency-bound kernels could then spread over more, narrower warps)?
    (2) FP64 FMA rate, as a candidate multiplier pipe."""
    import __graft_entry__ as entry

    mod = entry.load_package()
    lib = mod.lib()
    lib.ckzg_b200_selftest_mulbench.restype = C.c_int
    out = {"probe": "lanes_and_dfma", "partial_warps": [], "dfma": []}
    iters = 2000
    for active in (32, 16, 8, 4, 1):
        ms = C.c_float(0)
        rc = lib.ckzg_b200_selftest_mulbench(1 | (active << 8), iters, 148, 128, C.byref(ms))
        out["partial_warps"].append({"active_lanes": active, "ms": ms.value, "us_per_dependent_product": ms.value * 1e3 / iters, "rc": rc})
    for bps, threads in ((4, 256), (8, 256)):
        ms = C.c_float(0)
        rc = lib.ckzg_b200_selftest_mulbench(1 << 16, iters, 148 * bps, threads, C.byref(ms))
        fmas = 148 * bps * threads * iters * 8
        out["dfma"].append({"blocks_per_sm": bps, "threads": threads, "ms": ms.value, "dfma_per_s": fmas / (ms.value * 1e-3), "dfma_per_clk_per_sm_at_1965": fmas / (ms.value * 1e-3) / 148 / 1.965e9, "rc": rc})
    print(json.dumps(out))


def placement():
    """Where do the warps of the stage-1 kernels run?  Arms the engine's placement probe, runs one
    device-resident verify (n = 4096) per stage-1 arrangement, and reports for every kernel the number of
    SMs used, the spread over hardware warp slots (slot % 4 = sub-partition) and how many warps share a
    sub-partition with another recorded warp."""
    import collections

    import torch

    import __graft_entry__ as entry
    import bench

    mod = entry.load_package()
    lib = mod.lib()
    ts = mod.load_trusted_setup()
    n = int(os.environ.get("PROBE_N", "4096"))
    dev = torch.from_numpy(bench.synth_blobs(n, 7)).cuda()
    cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
    mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
    cap = 1 << 16
    buf = torch.zeros(cap + 2, dtype=torch.int32, device="cuda")
    buf[1] = cap
    for _ in range(2):
        assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)
    torch.cuda.synchronize()
    lib.ckzg_b200_debug_placement.argtypes = [C.c_void_p]
    assert lib.ckzg_b200_debug_placement(buf.data_ptr()) == 0
    assert mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)
    torch.cuda.synchronize()
    lib.ckzg_b200_debug_placement(None)
    h = buf.cpu().numpy().astype("uint32")
    cnt = int(h[0])
    recs = h[2 : 2 + min(cnt, cap)]
    names = {1: "blob_challenge", 2: "g1_validate2", 3: "hash+validate(fused)"}
    per = collections.defaultdict(list)
    occ = collections.Counter()
    for r in recs:
        r = int(r)
        k, w, sm, slot = r >> 28, (r >> 24) & 15, (r >> 8) & 0xffff, r & 0xff
        per[k].append((w, sm, slot))
        occ[(sm, slot % 4)] += 1
    out = {"probe": "placement", "n": n, "stage1": os.environ.get("CKZG_B200_STAGE1", "1"), "records": cnt, "kernels": {}}
    for k, lst in per.items():
        sms = {sm for _, sm, _ in lst}
        out["kernels"][names.get(k, str(k))] = {
            "warps": len(lst), "sms_used": len(sms),
            "slot_mod4_by_warp_in_block": {str(w): dict(collections.Counter(slot % 4 for ww, _, slot in lst if ww == w)) for w in sorted({w for w, _, _ in lst})},
            "warps_sharing_a_subpartition": sum(1 for _, sm, slot in lst if occ[(sm, slot % 4)] > 1),
        }
    print(json.dumps(out))


def commit_variant():
    import numpy as np
    import torch

    import __graft_entry__ as entry

    mod = entry.load_package()
    ts = mod.load_trusted_setup()
    sys.path.insert(0, ROOT)
    import bench

    n = 1024
    blobs = torch.from_numpy(bench.synth_blobs(n, 99)).cuda()
    out = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        mod.blob_to_kzg_commitment_device(out.data_ptr(), blobs.data_ptr(), n, ts)
    mod.profile_enable(ts, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        mod.blob_to_kzg_commitment_device(out.data_ptr(), blobs.data_ptr(), n, ts)
    dt = (time.perf_counter() - t0) / 3
    prof = mod.profile_dump(ts)
    print(json.dumps({"probe": "commit_variant", "variant": os.environ.get("CKZG_B200_ACC_VARIANT", "3"), "blobs_per_s": n / dt, "ms_per_call": dt * 1e3,
                      "kernels_ms_per_call": {k: v[0] / 3 for k, v in prof["kernels"].items()}}))


def verify_modes():
    """verify n=4096 device-resident: wall time per call with stages concurrent (profile level 0/1) and
    serialised with per-kernel events (level 2); same with pinned host pointers."""
    import torch

    import __graft_entry__ as entry
    import bench

    mod = entry.load_package()
    ts = mod.load_trusted_setup()
    n = int(os.environ.get("PROBE_N", "4096"))
    host = torch.from_numpy(bench.synth_blobs(n, 7)).pin_memory()
    dev = host.cuda()
    cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
    mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
    hc, hp = cms.cpu().pin_memory(), prs.cpu().pin_memory()
    out = {"probe": "verify_modes", "n": n, "rlc": os.environ.get("CKZG_B200_RLC", "vmsm")}
    for name, level, fn in (
        ("device_concurrent", 0, lambda: mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)),
        ("device_level1", 1, lambda: mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)),
        ("device_serial_level2", 2, lambda: mod.verify_blob_kzg_proof_batch_device(dev.data_ptr(), cms.data_ptr(), prs.data_ptr(), n, ts)),
        ("host_concurrent", 0, lambda: mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)),
        ("host_level1", 1, lambda: mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)),
        ("host_serial_level2", 2, lambda: mod.verify_blob_kzg_proof_batch_host(host.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, ts)),
    ):
        mod.profile_enable(ts, level)
        for _ in range(3):
            assert fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(6):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        out[name + "_ms"] = round(best * 1e3, 3)
        if level == 1:  # stage boundaries of the concurrent arrangement (events on the call's stream)
            pr = mod.profile_dump(ts)
            out[name + "_stages_ms"] = {k: round(v[0] / max(1, pr["calls"]), 3) for k, v in pr["kernels"].items() if k != "begin"}
        if level == 2:
            pr = mod.profile_dump(ts)
            out[name + "_kernels_ms"] = {k: round(v[0] / max(1, pr["calls"]), 3) for k, v in pr["kernels"].items() if k != "begin"}
    mod.profile_enable(ts, 0)
    print(json.dumps(out))


def single_commit():
    import __graft_entry__ as entry
    import bench

    mod = entry.load_package()
    ts = mod.load_trusted_setup()
    blob = bench.synth_blobs(1, 5).tobytes()
    for _ in range(3):
        mod.blob_to_kzg_commitment(blob, ts)
    mod.profile_enable(ts, 2)
    t0 = time.perf_counter()
    for _ in range(10):
        mod.blob_to_kzg_commitment(blob, ts)
    dt = (time.perf_counter() - t0) / 10
    prof = mod.profile_dump(ts)
    print(json.dumps({"probe": "single_commit", "ms_per_call_wall": dt * 1e3, "device_ms_per_call": prof["call_ms"] / 10, "kernels_ms_per_call": {k: v[0] / 10 for k, v in prof["kernels"].items()}}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "modes":
        verify_modes()
    elif len(sys.argv) > 1 and sys.argv[1] == "placement":
        placement()
    elif len(sys.argv) > 1 and sys.argv[1] == "lanes":
        lanes_and_dfma()
    elif len(sys.argv) > 1 and sys.argv[1] == "single":
        single_commit()
    elif len(sys.argv) > 1 and sys.argv[1] == "commit":
        commit_variant()
    else:
        mulbench()
        subprocess.call([sys.executable, os.path.abspath(__file__), "modes"])
        subprocess.call([sys.executable, os.path.abspath(__file__), "single"])
        if os.environ.get("PROBE_COMMIT_VARIANTS"):  # settled in r01l; kept for re-measurement
            for v in ("3", "13", "4", "14"):
                env = dict(os.environ, CKZG_B200_ACC_VARIANT=v)
                subprocess.call([sys.executable, os.path.abspath(__file__), "commit"], env=env)
