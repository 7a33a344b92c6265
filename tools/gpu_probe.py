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
    n = 4096
    host = torch.from_numpy(bench.synth_blobs(n, 7)).pin_memory()
    dev = host.cuda()
    cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
    mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
    mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
    hc, hp = cms.cpu().pin_memory(), prs.cpu().pin_memory()
    out = {"probe": "verify_modes", "stage1_mode": os.environ.get("CKZG_B200_STAGE1_MODE", "1")}
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
