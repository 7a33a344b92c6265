#!/usr/bin/env python3
"""bench.py -- blobs/sec of the KZG hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, all host cores

Headline workload: verify_blob_kzg_proof_batch at the north-star batch size (4096 blobs per GPU), the
configuration BASELINE.json's metric and target are quoted on.  One "step" = one pass of the hot path over one
batch of synthetic blobs PLUS the collective that combines the ranks' verdicts.

  value ........ blobs/s with blobs, commitments and proofs already resident in HBM (device pointers into the
                 engine's C ABI); timed with CUDA events on the legacy default stream around the K steps -- the
                 engine's call streams are blocking streams, so those events bracket the engine's kernels AND the
                 NCCL all-reduce of each step; max over ranks
  e2e .......... the same step with HOST pointers (pinned): H2D of every input and D2H of the verdict inside the
                 timed region; e2e_pageable = plain pageable host memory (what Go slices / Python bytes are)
  commitment ... blob_to_kzg_commitment (the other function the metric names), batch 1024 per GPU, at every N
  configs ...... BASELINE configs[1..4] and the north-star batch (cells + proofs at 4096) at every N, each with its
                 collective inside the timed region
  roofline ..... the dominant kernel against the INTEGER-pipe roof (the path is integer-bound), peak =
                 the Montgomery-multiplier microbenchmark run live in this process; HBM as the secondary entry
  cpu_baseline . the unmodified reference (oracle/_ref) on 1 host core, bounded sample (rank 0, N=1)

Multi-GPU (--gpus N under torchrun): weak scaling, each rank verifies its own 4096-blob batch and the verdicts are
combined with one NCCL all-reduce(MIN) per step (c-kzg-4844_b200/parallel.py).  The single-challenge sharding of ONE
global batch (`sharded_single_challenge`) and the in-library multi-device context (`inlib_multi_device`, rank 0 drives
all N GPUs through the frozen-API entry point) are measured beside it at N > 1.
Inputs are larger than L2 (512 MiB of blobs per step vs 126 MB), so no explicit flush is needed.
"""
import os as _os

# four hardware work queues per device (must be set before CUDA is initialised; see c-kzg-4844_b200/csrc/api.cu)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "4")
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOB = 131072
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
METRIC = "blobs/sec: blob_to_kzg_commitment & verify_blob_kzg_proof_batch @1/2/4/8 GPU"
DTYPE = "u32 limbs (381/255-bit Montgomery integers)"

# algorithmic HBM bytes per blob for each kernel of the verify step (DESIGN.md "Roofline accounting")
ALGO_BYTES_PER_BLOB = {
    "evaluate": 131072 + 12 * 32 + 64,     # blob in, powers of z in, z||y out (roots of unity are L2-resident constants)
    "hash+validate": 131072 + 48 + 64 + 2 * (48 + 96) + 2 * 18 * 192,  # fused stage-1 kernel: blob + points in, z, affine points and table columns out
    "vmsm_accumulate": 2 * 3 * 32 * (4 + 192) + 192 * 3 * 32 // 8,  # ~96 digit entries per blob: index + 192-byte table point in, one partial per 8 entries out
    "vmsm_sort": 3 * 32 * 2 + 3 * 32 * 4 * 2,
    "vmsm_reduce": 192 * 3 * 32 // 8,
    "transcript(d2h,host_sha)": 160,
    "blob_challenge": 131072 + 48 + 64,    # blob + commitment in, z out
    "g1_validate": 2 * (48 + 96) + 2 * 18 * 192,  # two points in, two affine points + their 18 table levels out
    "rlc_points": 3 * 96 + 64 + 3 * 192,   # 3 bases + 2 scalars in, 3 XYZZ out
    "rlc_scalars": 64 + 96,
    "g1_sum": 3 * 192,
    "pairing_check": 0,
    "msm_direct": 131072 + 77824 * 96 + 8 * 192,  # blob in, one 96-byte table entry per (scalar, window), 8 partial sums out
}
# measured DRAM traffic per blob (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture)
MEASURED_TRAFFIC_PER_BLOB = {"hash+validate": 135_400}  # profiles/ncu_r01t_summary.txt: 539 MB + 15 MB per 4096 blobs

# algorithmic 32x32->64 multiply-accumulates per blob (SURVEY.md section 8(d) convention: Fp mul = 300, Fr mul = 136)
ALGO_MAC_PER_BLOB = {
    "evaluate": (4096 + 2 * 4095) * 136,   # tree evaluation: one product per leaf, two per node
    "g1_validate": 2 * 2200 * 300,
    "hash+validate": 2 * 2200 * 300,       # + 2050 SHA-256 blocks per blob, which are not multiplications
    "rlc_points": 3 * 3600 * 300,
    "vmsm_accumulate": 3 * 32 * 14 * 300,  # ~32 non-zero signed bytes per scalar, three scalars per blob, 14 products per XYZZ addition
    "msm_direct": 77824 * 10 * 300,        # 4096 x 19 windows of mixed XYZZ additions (10 products each)
}
# per CALL, not per blob: one pairing check = two Miller loops sharing their squarings + one final exponentiation,
# ~19,000 dependent Fp products on ONE CTA -- a latency chain, so its pipe fraction is ~0 by construction
ALGO_MAC_PER_CALL = {"pairing_check": 19000 * 300}
# per-blob MACs of the reference algorithm for the other configs (SURVEY 8d "Work per unit")
REF_MAC_PER_BLOB = {"commitment": 4.3e8, "blob_verify": 7.1e6, "cells_and_proofs": 1.1e9, "recover": 1.15e9, "cell_verify": 7.9e7}


def workload_config(n, world, sharded=False):
    """The `config` object of BOTH arms (the reference arm measures the same workload on the host cores)."""
    return {
        "workload": "verify_blob_kzg_proof_batch n=%d per call, one call per GPU per step (north_star batch)" % n,
        "blobs_per_gpu": n,
        "parallelism": ("sharded global batch x%d: all-gather(z||y) + all-gather(partials)" % world) if sharded else ("replicas x%d + 1 all-reduce(MIN) per step" % world if world > 1 else "single GPU"),
        "l2": "inputs (%.0f MiB/step) exceed the 126 MB L2; no explicit flush" % (n * BLOB / 2**20),
    }


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (profiling recipe's clocks line)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # samples taken inside the timed region [t0, t1]; if the region was shorter than the sampling
        # period fall back to the samples of the surrounding warm-up (same load)
        rows = [r[1:] for r in self.rows if len(r) >= 8 and (t0 is None or (t0 - 0.02 <= r[0] <= t1 + 0.02))]
        window = "timed region"
        if not rows:
            rows, window = [r[1:] for r in self.rows if len(r) >= 8], "warm-up + timed region"
        sm = sorted(int(float(r[0])) for r in rows if r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "window": window}


def synth_blobs(n, seed):
    """n random canonical blobs as one uint8 array: 32 random bytes per element, top byte drawn from
    [0, 0x72] so every element is < r (0x73ed...) and spans the full 255-bit range."""
    import numpy as np

    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n * 4096, 32), dtype=np.uint8)
    a[:, 0] = rng.integers(0, 0x73, size=n * 4096, dtype=np.uint8)
    return a.reshape(-1)


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference CPU implementation on all host cores
# ------------------------------------------------------------------------------------------------
_REF_INPUT = None  # (blobs, commitments, proofs) built once in the parent; forked workers share the pages


def _ref_worker(args):
    wid, steps, warmup = args
    from oracle import ref_lib

    k = ref_lib.CKZG()
    B, Cc, P = _REF_INPUT
    for _ in range(warmup):
        assert k.verify_blob_kzg_proof_batch(B, Cc, P)
    t0 = time.perf_counter()
    for _ in range(steps):
        assert k.verify_blob_kzg_proof_batch(B, Cc, P)
    return (len(Cc) // 48) * steps, time.perf_counter() - t0


def run_reference(args):
    global _REF_INPUT
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import ref_lib

    if not os.path.exists(ref_lib.REF_SO):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libckzg_ref.so missing (build with oracle/build_ref.sh)"}))
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # the same call as our arm: n = 4096 blobs per call, one call per step on EVERY host core at once (one process per
    # core, as the reference's own parallel benchmarks run it: bindings/go/main_test.go:957-970).  ~5 s per call per core.
    n = args.blobs
    n_distinct = 32
    tile = max(1, n // n_distinct)
    steps, warmup = args.steps, args.warmup
    k = ref_lib.CKZG()
    blobs = synth_blobs(n_distinct, 4844 + 1000).tobytes()
    bl = [blobs[BLOB * i : BLOB * (i + 1)] for i in range(n_distinct)]
    cms = [k.blob_to_kzg_commitment(b) for b in bl]
    prs = [k.compute_blob_kzg_proof(b, c) for b, c in zip(bl, cms)]
    k.close()
    _REF_INPUT = (blobs * tile, b"".join(cms) * tile, b"".join(prs) * tile)
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, [(w, steps, warmup) for w in range(cores)])
    total = sum(r[0] for r in res)
    tmax = max(r[1] for r in res)
    value = total / tmax
    sample = "each of %d processes (one per host core): verify_blob_kzg_proof_batch n=%d (%d distinct synthetic blobs tiled x%d), %d timed calls after %d warm-up" % (
        cores, n_distinct * tile, n_distinct, tile, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "blobs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * tmax / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": workload_config(n, args.gpus),
        "cpu_baseline": {"value": value, "unit": "blobs/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "blobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def cpu_baseline_single_core(blobs, cms, prs, n, seconds_budget=25.0):
    from oracle import ref_lib

    if not os.path.exists(ref_lib.REF_SO):
        return None
    k = ref_lib.CKZG()
    m = min(n, 4096)
    B, Cc, P = blobs[: BLOB * m], cms[: 48 * m], prs[: 48 * m]
    best, calls, t_start = None, 0, time.perf_counter()
    while calls < 2 and (time.perf_counter() - t_start) < seconds_budget:
        t0 = time.perf_counter()
        ok = k.verify_blob_kzg_proof_batch(B, Cc, P)
        dt = time.perf_counter() - t0
        assert ok
        best = dt if best is None else min(best, dt)
        calls += 1
    k.close()
    return {"value": m / best, "unit": "blobs/s", "cores": 1, "kind": "reference",
            "sample": "oracle/_ref (unmodified reference, blst ADX path) verify_blob_kzg_proof_batch n=%d, best of %d calls on one core" % (m, calls)}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    os.environ["CKZG_B200_DEVICE"] = str(local_rank)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    # host threads per rank (staging memcpy, verify_cell sub-batch hashes): the ranks share the box's cores
    os.environ.setdefault("CKZG_B200_HOST_THREADS", str(max(2, min(8, ncpu // world))))
    mod = entry.load_package()
    par = __import__("importlib").import_module("ckzg_b200.parallel")
    ts = mod.load_trusted_setup()
    n = args.blobs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def reduce_min(vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return [float(x) for x in t]

    def all_ok(flag):
        """the step's collective: one all-reduce(MIN) of the rank's verdict / status, then the host reads it"""
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(step_fn, steps, warmup, sample_clocks=False):
        """W warm-up steps, barrier + synchronize, K steps between two CUDA events on the legacy default stream
        (blocking engine streams and NCCL's stream are ordered against it), synchronize + barrier.  Returns
        (device ms per step, wall ms per step) as the max over ranks, and the clock samples of the timed region."""
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
            t_wait = time.perf_counter()
            while not sampler.rows and time.perf_counter() - t_wait < 3.0:  # nvidia-smi needs ~0.2 s to print its first row
                time.sleep(0.01)
        for _ in range(warmup):
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        dev_ms, wall_ms = reduce_max([e0.elapsed_time(e1) / steps, 1000.0 * (t1 - t0) / steps])
        return dev_ms, wall_ms, (sampler.stop(t0, t1) if sampler else None)

    # ---- synthetic inputs; commitments and proofs produced by the (parity-tested) engine itself ----
    host_blobs = torch.from_numpy(synth_blobs(n, 4844 + rank)).pin_memory()
    d_blobs = host_blobs.to(dev, non_blocking=False)
    d_cms = torch.empty(48 * n, dtype=torch.uint8, device=dev)
    d_prs = torch.empty(48 * n, dtype=torch.uint8, device=dev)
    mod.blob_to_kzg_commitment_device(d_cms.data_ptr(), d_blobs.data_ptr(), n, ts)
    mod.compute_blob_kzg_proof_device(d_prs.data_ptr(), d_blobs.data_ptr(), d_cms.data_ptr(), n, ts)
    host_cms, host_prs = d_cms.cpu().pin_memory(), d_prs.cpu().pin_memory()

    def verify_dev():
        return mod.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_cms.data_ptr(), d_prs.data_ptr(), n, ts)

    def verify_host():
        return mod.verify_blob_kzg_proof_batch_host(host_blobs.data_ptr(), host_cms.data_ptr(), host_prs.data_ptr(), n, ts)

    def step_of(fn):
        def step():
            assert all_ok(fn()), "verification of a valid synthetic batch failed"
        return step

    # negative control: two swapped proofs must be rejected (no synchronize needed: blocking call streams)
    bad = d_prs.clone()
    bad[0:48], bad[48:96] = d_prs[48:96].clone(), d_prs[0:48].clone()
    assert not mod.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_cms.data_ptr(), bad.data_ptr(), n, ts), "negative control accepted"
    del bad

    # ---- headline: device-resident step incl. the collective.  The timed steps run UNPROFILED (a call under level-1
    # profiling takes streams of its own instead of the context's pooled ones and records timing events: ~0.3 ms per
    # call on one GPU, 3 ms with two processes creating and destroying streams side by side, profiles/bench_R3h.log);
    # the engine-internal stage marks come from a short profiled run right after ----
    l0 = ts.launch_count()
    ms_per_step, wall_ms, clocks = timed(step_of(verify_dev), args.steps, args.warmup, sample_clocks=True)
    launches = (ts.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    mod.profile_enable(ts, 1)
    for _ in range(2):  # the first profiled calls create their streams
        assert verify_dev()
    mod.profile_enable(ts, 1)  # resets the counters
    for _ in range(5):
        assert verify_dev()
    prof1 = mod.profile_dump(ts)
    mod.profile_enable(ts, 0)
    engine_ms = prof1["call_ms"] / max(1, prof1["calls"])
    value = world * n / (ms_per_step / 1000.0)

    # ---- e2e: the same step through HOST pointers (pinned), and through pageable memory ----
    e_ms, e_wall, _ = timed(step_of(verify_host), args.steps, max(3, args.warmup - 1))
    e2e_value = world * n / (e_ms / 1000.0)
    # where the end-to-end call's time goes: a separate short run with the engine's stage marks (timing events on the
    # call's streams; kept out of the timed run above), completion times relative to the call's first event
    mod.profile_enable(ts, 1)
    for _ in range(2):  # profiled calls take streams of their own: the first ones pay their creation
        verify_host()
    mod.profile_enable(ts, 1)  # resets the counters
    for _ in range(3):
        verify_host()
    prof_e = mod.profile_dump(ts)
    mod.profile_enable(ts, 0)
    e2e_stages = {k: round(v[0] / max(1, v[1]), 3) for k, v in prof_e.get("kernels", {}).items() if k.startswith("stage:") or k.startswith("host:")}
    e2e_stages["engine_ms_per_call"] = round(prof_e["call_ms"] / max(1, prof_e["calls"]), 3)
    pg_blobs = np.array(host_blobs.numpy(), copy=True)  # ordinary pageable memory: what a Go slice or Python bytes is
    pg_cms, pg_prs = np.array(host_cms.numpy(), copy=True), np.array(host_prs.numpy(), copy=True)

    def verify_pageable():
        return mod.verify_blob_kzg_proof_batch_host(pg_blobs.ctypes.data, pg_cms.ctypes.data, pg_prs.ctypes.data, n, ts)

    p_ms, _, _ = timed(step_of(verify_pageable), max(3, args.steps // 2), 3)
    e2e_pageable = {"value": world * n / (p_ms / 1000.0), "unit": "blobs/s", "ms_per_step": p_ms, "vs_pinned": e_ms / p_ms,
                    "note": "caller buffers in pageable host memory; staged through a pinned ring by %s host threads per rank (csrc/hostpool.h)" % os.environ["CKZG_B200_HOST_THREADS"]}
    del pg_blobs

    # ---- per-rank H2D bandwidth with all ranks copying at once (what bounds e2e as N grows) ----
    def h2d_probe():
        d_blobs.copy_(host_blobs, non_blocking=True)
    hp_ms, _, _ = timed(h2d_probe, 3, 1)
    h2d = {"GBps_per_rank_slowest": n * BLOB / (hp_ms * 1e-3) / 1e9, "concurrent_ranks": world, "bytes": n * BLOB,
           "note": "pinned host -> device copy of one step's blobs on every rank at once, max time over ranks"}
    d_blobs.copy_(host_blobs)
    torch.cuda.synchronize()

    # ---- kernel breakdown: per-kernel events (level 2, stages serialised on one stream); clocks warm, 5 calls ----
    for _ in range(3):
        verify_dev()
    mod.profile_enable(ts, 2)
    for _ in range(2):  # as above: the first profiled calls create their streams
        verify_dev()
    mod.profile_enable(ts, 2)
    prof_steps = 5
    for _ in range(prof_steps):
        verify_dev()
    prof = mod.profile_dump(ts)
    mod.profile_enable(ts, 0)
    kern = {k: v for k, v in prof["kernels"].items() if k not in ("begin", "end")}
    total_ms = sum(v[0] for v in kern.values()) or 1.0

    # ---- integer-pipe roof, measured now: the Montgomery multiplier microbenchmark (selftest_mulbench) ----
    mul_iters, mul_blocks, mul_threads, mul_ilp = 2000, 148 * 8, 256, 2
    mod.mulbench(mul_ilp, 200, mul_blocks, mul_threads)
    mb_ms = min(mod.mulbench(mul_ilp, mul_iters, mul_blocks, mul_threads) for _ in range(3))
    fp_mul_per_s = mul_blocks * mul_threads * mul_ilp * mul_iters / (mb_ms * 1e-3)
    peak_mac = fp_mul_per_s * 300.0
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965
    planning_mac = 148 * 64 * sm_mhz * 1e6

    hbm_peak, hbm_src = load_peaks()
    macs = {k: ALGO_MAC_PER_BLOB[k] * n / (kern[k][0] / prof_steps * 1e-3) for k in ALGO_MAC_PER_BLOB if k in kern and kern[k][0] > 0}
    for k, per_call in ALGO_MAC_PER_CALL.items():
        if k in kern and kern[k][0] > 0:
            macs[k] = per_call / (kern[k][0] / kern[k][1] * 1e-3)
    dom = max(kern, key=lambda k: kern[k][0])                      # largest share of the step
    dom_mac = max(macs, key=lambda k: kern[k][0]) if macs else dom  # largest MAC-carrying kernel
    dom_ms = kern[dom][0] / kern[dom][1]
    units_per_launch = n / (kern[dom][1] / prof_steps)
    achieved_mac = macs.get(dom, 0.0)
    algo_bytes = ALGO_BYTES_PER_BLOB.get(dom, 0) * units_per_launch
    step_mac = REF_MAC_PER_BLOB["blob_verify"] * n / (ms_per_step * 1e-3)  # the TIMED step (collective and host read included)
    roofline = {
        "bound": "int_pipe", "kernel": dom, "achieved": achieved_mac / 1e12, "peak": peak_mac / 1e12, "unit": "TMAC/s (32x32->64 multiply-accumulates; Fp product = 300, Fr product = 136: SURVEY 8d)",
        "frac": achieved_mac / peak_mac,
        "peak_source": "measured in this run: Fp Montgomery-multiplier microbenchmark (ckzg_b200_selftest_mulbench, %d x %d threads, ilp %d) = %.2f G Fp products/s" % (mul_blocks, mul_threads, mul_ilp, fp_mul_per_s / 1e9),
        "frac_of_planning_peak": achieved_mac / planning_mac, "planning_peak": "148 SMs x 64 MAC/clk x %d MHz (SURVEY 8d)" % sm_mhz,
        "share_of_step": kern[dom][0] / total_ms, "ms_per_launch": dom_ms, "algorithmic_mac_per_launch": ALGO_MAC_PER_CALL.get(dom, ALGO_MAC_PER_BLOB.get(dom, 0) * units_per_launch),
        "traffic": (MEASURED_TRAFFIC_PER_BLOB[dom] * units_per_launch) if dom in MEASURED_TRAFFIC_PER_BLOB else None,
        "note": ("%s also runs the 2050-block SHA-256 chain of every blob (ALU/FMA-pipe additions and rotations, no multiplications)" % dom) if dom == "hash+validate"
                else ("one pairing check per call on ONE 64-thread CTA: a chain of dependent tower operations, latency-bound by construction" if dom == "pairing_check" else ""),
        "hbm": {"bound": "hbm", "achieved": algo_bytes / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": algo_bytes / (dom_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": hbm_src, "algorithmic_bytes_per_blob": ALGO_BYTES_PER_BLOB.get(dom, 0)},
        "whole_step": {"mac_per_blob_reference_algorithm": REF_MAC_PER_BLOB["blob_verify"], "achieved": step_mac / 1e12, "frac": step_mac / peak_mac, "frac_of_planning_peak": step_mac / planning_mac},
        "largest_mac_kernel": {"kernel": dom_mac, "frac": macs.get(dom_mac, 0) / peak_mac, "share_of_step": kern[dom_mac][0] / total_ms} if macs else None,
        "per_kernel": {
            k: {"ms_per_step": round(v[0] / prof_steps, 4), "share": round(v[0] / total_ms, 4),
                "int_pipe_frac": round(macs[k] / peak_mac, 4) if k in macs else None,
                "algorithmic_GBps": round(ALGO_BYTES_PER_BLOB.get(k, 0) * n / (v[0] / prof_steps * 1e-3) / 1e9, 2) if v[0] > 0 else None}
            for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])
        },
    }

    # ---- blob_to_kzg_commitment (the other function the metric names): batch 1024 per GPU, at every N ----
    mc = min(n, 1024)
    d_out = torch.empty(48 * mc, dtype=torch.uint8, device=dev)
    h_out = torch.empty(48 * mc, dtype=torch.uint8).pin_memory()
    gathered = [torch.empty_like(d_out) for _ in range(world)] if world > 1 else None

    def commit_dev():
        mod.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), mc, ts)
        if world > 1:
            dist.all_gather(gathered, d_out)  # the step's collective: every rank ends with all commitments
        d_out[:1].cpu()

    def commit_host():
        mod.blob_to_kzg_commitment_batch_host(h_out.data_ptr(), host_blobs.data_ptr(), mc, ts)
        all_ok(True)

    c_ms, _, _ = timed(commit_dev, 5, 3)
    ce_ms, _, _ = timed(commit_host, 5, 2)
    mod.profile_enable(ts, 2)
    for _ in range(3):
        mod.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), mc, ts)
    pc = mod.profile_dump(ts)
    mod.profile_enable(ts, 0)
    md_ms = pc["kernels"].get("msm_direct", [0, 1])[0] / 3
    commitment = {
        "value": world * mc / (c_ms * 1e-3), "unit": "blobs/s", "batch_per_gpu": mc, "ms_per_step": c_ms,
        "e2e": {"value": world * mc / (ce_ms * 1e-3), "unit": "blobs/s", "h2d_bytes_per_step": mc * BLOB, "d2h_bytes_per_step": mc * 48},
        "collective": "all-gather of the %d x 48-byte commitments" % mc if world > 1 else "none (single GPU)",
        "int_pipe": {"kernel": "msm_direct", "ms_per_launch": md_ms, "achieved": (ALGO_MAC_PER_BLOB["msm_direct"] * mc / (md_ms * 1e-3) / 1e12) if md_ms else None,
                     "peak": peak_mac / 1e12, "frac": (ALGO_MAC_PER_BLOB["msm_direct"] * mc / (md_ms * 1e-3) / peak_mac) if md_ms else None,
                     "hbm_GBps_algorithmic": (ALGO_BYTES_PER_BLOB["msm_direct"] * mc / (md_ms * 1e-3) / 1e9) if md_ms else None},
        # the reference's Pippenger needs 1.42 M Fp products per blob, the direct table 0.78 M: > 1 means "more than the
        # measured multiplier could deliver on the reference's algorithm"
        "reference_algorithm_equivalent_frac_of_int_peak": REF_MAC_PER_BLOB["commitment"] * mc / (c_ms * 1e-3) / peak_mac,
    }

    # ---- BASELINE configs[2..4] + the north-star batch, on every rank, each with its collective in the timed region ----
    configs = {}
    if not args.no_extra:
        m7 = min(n, 256)
        d_cells = torch.empty(m7 * 2 * BLOB, dtype=torch.uint8, device=dev)
        d_cprf = torch.empty(m7 * 128 * 48, dtype=torch.uint8, device=dev)

        def cells_dev():
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), d_cprf.data_ptr(), d_blobs.data_ptr(), m7, ts)
            all_ok(True)

        t_ms, _, _ = timed(cells_dev, 3, 2)
        h_cells = torch.empty(m7 * 2 * BLOB, dtype=torch.uint8).pin_memory()
        h_cprf = torch.empty(m7 * 128 * 48, dtype=torch.uint8).pin_memory()

        def cells_host():
            mod.compute_cells_and_kzg_proofs_host(h_cells.data_ptr(), h_cprf.data_ptr(), host_blobs.data_ptr(), m7, ts)
            all_ok(True)

        te_ms, _, _ = timed(cells_host, 3, 1)
        configs["compute_cells_and_kzg_proofs_256_per_gpu"] = {
            "baseline_config": "configs[2]: compute_cells_and_kzg_proofs, 256 blobs per GPU", "value": world * m7 / (t_ms * 1e-3), "unit": "blobs/s", "ms_per_step": t_ms,
            "e2e": world * m7 / (te_ms * 1e-3), "collective": "all-reduce(MIN) of the status",
            "reference_algorithm_frac_of_int_peak": REF_MAC_PER_BLOB["cells_and_proofs"] * m7 / (t_ms * 1e-3) / peak_mac}
        mod.profile_enable(ts, 2)
        for _ in range(2):
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), d_cprf.data_ptr(), d_blobs.data_ptr(), m7, ts)
        p7 = mod.profile_dump(ts)
        mod.profile_enable(ts, 0)
        configs["compute_cells_and_kzg_proofs_256_per_gpu"]["kernels_ms"] = {k: round(v[0] / 2, 3) for k, v in p7["kernels"].items() if k not in ("begin", "end")}

        # north-star batch: 4096 blobs per GPU in one call (outputs 1 GiB of cells + 25 MB of proofs per GPU)
        if n >= 4096:
            big_cells = torch.empty(n * 2 * BLOB, dtype=torch.uint8, device=dev)
            big_prf = torch.empty(n * 128 * 48, dtype=torch.uint8, device=dev)

            def cells_big():
                mod.compute_cells_and_kzg_proofs_device(big_cells.data_ptr(), big_prf.data_ptr(), d_blobs.data_ptr(), n, ts)
                all_ok(True)

            tb_ms, _, _ = timed(cells_big, 2, 1)
            configs["compute_cells_and_kzg_proofs_4096_per_gpu"] = {
                "baseline_config": "north_star: compute_cells_and_kzg_proofs at batch=4096 per GPU", "value": world * n / (tb_ms * 1e-3), "unit": "blobs/s", "ms_per_step": tb_ms,
                "collective": "all-reduce(MIN) of the status", "reference_algorithm_frac_of_int_peak": REF_MAC_PER_BLOB["cells_and_proofs"] * n / (tb_ms * 1e-3) / peak_mac}
            del big_cells, big_prf
            torch.cuda.empty_cache()

        # configs[3]: recover from the even-indexed cells (50 % missing), 256 blobs per GPU (= 1024 blobs on 4 GPUs)
        idx = list(range(0, 128, 2)) * m7
        given = d_cells.view(m7, 128, 2048)[:, 0::2, :].contiguous()
        rec_c = torch.empty(m7 * 2 * BLOB, dtype=torch.uint8, device=dev)
        rec_p = torch.empty(m7 * 128 * 48, dtype=torch.uint8, device=dev)
        mod.recover_cells_and_kzg_proofs_device(rec_c.data_ptr(), rec_p.data_ptr(), idx, given.data_ptr(), 64, m7, ts)
        assert torch.equal(rec_c, d_cells) and torch.equal(rec_p, d_cprf), "recover != compute"

        def recover_dev():
            mod.recover_cells_and_kzg_proofs_device(rec_c.data_ptr(), rec_p.data_ptr(), idx, given.data_ptr(), 64, m7, ts)
            all_ok(True)

        tr_ms, _, _ = timed(recover_dev, 3, 1)
        configs["recover_cells_and_kzg_proofs_256_per_gpu"] = {
            "baseline_config": "configs[3]: recover_cells_and_kzg_proofs, 50% missing (even cells given), 256 blobs per GPU (1024 blobs at 4 GPUs)",
            "value": world * m7 / (tr_ms * 1e-3), "unit": "blobs/s", "ms_per_step": tr_ms, "collective": "all-reduce(MIN) of the status",
            "reference_algorithm_frac_of_int_peak": REF_MAC_PER_BLOB["recover"] * m7 / (tr_ms * 1e-3) / peak_mac}
        del rec_c, rec_p, given

        # configs[4]: verify_cell_kzg_proof_batch, 512 blobs x 128 cells per GPU (= 4096 x 128 on 8 GPUs), row-major by blob
        # (bindings/go/main_test.go:1016-1028), HOST pointers: 141 MB of input per GPU per step
        nb = min(n, 512)
        v_cells = torch.empty(nb * 2 * BLOB, dtype=torch.uint8).pin_memory()
        v_prf = torch.empty(nb * 128 * 48, dtype=torch.uint8).pin_memory()
        for off in range(0, nb, m7):
            mm = min(m7, nb - off)
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), d_cprf.data_ptr(), d_blobs.data_ptr() + off * BLOB, mm, ts)
            v_cells[off * 2 * BLOB : (off + mm) * 2 * BLOB].copy_(d_cells[: mm * 2 * BLOB])
            v_prf[off * 6144 : (off + mm) * 6144].copy_(d_cprf[: mm * 6144])
        v_cm = host_cms.view(-1, 48)[:nb].repeat_interleave(128, dim=0).contiguous().pin_memory()
        v_idx = (ctypes.c_uint64 * (nb * 128))(*([k for _ in range(nb) for k in range(128)]))

        def vcells():
            assert all_ok(mod.verify_cell_kzg_proof_batch_ptr(v_cm.data_ptr(), v_idx, v_cells.data_ptr(), v_prf.data_ptr(), nb * 128, ts)), "cell verification failed"

        tv_ms, _, _ = timed(vcells, 3, 1)
        v_prf2 = v_prf.clone()
        ci = nb * 128 - 5  # a proof of the last blob replaced by its neighbour's
        v_prf2[48 * ci : 48 * ci + 48] = v_prf[48 * (ci + 1) : 48 * (ci + 2)]
        assert not mod.verify_cell_kzg_proof_batch_ptr(v_cm.data_ptr(), v_idx, v_cells.data_ptr(), v_prf2.data_ptr(), nb * 128, ts), "cell-proof negative control accepted"
        configs["verify_cell_kzg_proof_batch_512x128_per_gpu"] = {
            "baseline_config": "configs[4]: verify_cell_kzg_proof_batch, 512 blobs x 128 cells per GPU (4096 x 128 at 8 GPUs), host pointers",
            "value": world * nb / (tv_ms * 1e-3), "unit": "blobs/s", "cells_per_s": world * nb * 128 / (tv_ms * 1e-3), "ms_per_step": tv_ms,
            "h2d_bytes_per_step": nb * 128 * (2048 + 48 + 48), "collective": "all-reduce(MIN) of the verdicts",
            "sub_batches": "each call is verified as independent sub-batches (own Fiat-Shamir challenge, one host thread each: CKZG_B200_HOST_THREADS=%s), verdicts AND-ed" % os.environ["CKZG_B200_HOST_THREADS"],
            "reference_algorithm_frac_of_int_peak": REF_MAC_PER_BLOB["cell_verify"] * nb / (tv_ms * 1e-3) / peak_mac}
        del v_prf2

        # configs[1]: verify_blob_kzg_proof_batch n = 64 in one call (latency-bound case)
        def v64():
            assert all_ok(mod.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_cms.data_ptr(), d_prs.data_ptr(), 64, ts))

        t64, _, _ = timed(v64, 10, 3)
        configs["verify_blob_kzg_proof_batch_n64"] = {"baseline_config": "configs[1]: verify_blob_kzg_proof_batch n=64 per GPU per call", "value": world * 64 / (t64 * 1e-3), "unit": "blobs/s", "ms_per_step": t64}

    # ---- N > 1: ONE global batch of N x n blobs with a single challenge (exact reference semantics for the whole
    #      batch): per-blob stage per rank, all-gather(z||y), partial bucket MSMs with r^(first+i), all-gather(partials) ----
    multi = {}
    if world > 1 and not args.no_extra:
        gc = [torch.empty_like(d_cms) for _ in range(world)]
        gp = [torch.empty_like(d_prs) for _ in range(world)]
        dist.all_gather(gc, d_cms)
        dist.all_gather(gp, d_prs)
        all_cms = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gc)
        all_prs = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gp)
        state = {}

        def s1():
            state["sh"] = mod.VerifyShard(d_blobs.data_ptr(), d_cms.data_ptr(), d_prs.data_ptr(), n, ts)
            return state["sh"].zy

        def s2(tuples, nt, first, nl):
            try:
                return state["sh"].stage2(tuples, nt, first)
            finally:
                state["sh"].close()

        def sharded_step():
            assert par.verify_batch_sharded(s1, s2, lambda parts, nr: mod.verify_shard_finish(parts, nr, ts), all_cms, all_prs, world * n, dev, pack=mod.pack_verify_tuples)

        ts_ms, _, _ = timed(sharded_step, 5, 2)
        multi["sharded_single_challenge"] = {
            "value": world * n / (ts_ms * 1e-3), "unit": "blobs/s", "ms_per_step": ts_ms, "global_batch": world * n,
            "collectives": "all-gather(z||y, 64 B/blob) + all-gather(partials, 384 B/rank); every rank hashes the %d-byte transcript and runs the pairing" % (32 + 160 * world * n)}

    if args.sharded and world > 1:
        value = multi["sharded_single_challenge"]["value"]
        ms_per_step = multi["sharded_single_challenge"]["ms_per_step"]

    # ---- N > 1: the in-library multi-device context: rank 0 alone drives all N GPUs through the frozen-API entry point
    #      (host pointers, pageable memory); the other ranks wait at the barrier with their GPUs idle ----
    if world > 1 and not args.no_extra:
        barrier()
        if rank == 0:
            try:
                os.environ["CKZG_B200_DEVICES"] = ",".join(str(i) for i in range(world))
                os.environ["CKZG_B200_HOST_THREADS"] = str(max(2, min(8, ncpu // world)))
                ts_all = mod.load_trusted_setup()
                del os.environ["CKZG_B200_DEVICES"]
                reps = world
                big_b = np.tile(host_blobs.numpy(), reps)
                big_c, big_p = np.tile(host_cms.numpy(), reps), np.tile(host_prs.numpy(), reps)

                def inlib(bp, cp, pp):
                    call = lambda: mod.verify_blob_kzg_proof_batch_host(bp, cp, pp, reps * n, ts_all)
                    assert call()
                    t0 = time.perf_counter()
                    for _ in range(3):
                        assert call()
                    return (time.perf_counter() - t0) / 3

                dt = inlib(big_b.ctypes.data, big_c.ctypes.data, big_p.ctypes.data)
                pin_b, pin_c, pin_p = torch.from_numpy(big_b).pin_memory(), torch.from_numpy(big_c).pin_memory(), torch.from_numpy(big_p).pin_memory()
                dt_pin = inlib(pin_b.data_ptr(), pin_c.data_ptr(), pin_p.data_ptr())
                multi["inlib_multi_device"] = {
                    "value": reps * n / dt_pin, "unit": "blobs/s", "ms_per_call": dt_pin * 1e3, "blobs_per_call": reps * n, "devices": world, "memory": "pinned host",
                    "pageable": {"value": reps * n / dt, "ms_per_call": dt * 1e3, "note": "staged through pinned rings by %s host threads per device" % os.environ["CKZG_B200_HOST_THREADS"]},
                    "what": "ONE verify_blob_kzg_proof_batch call (frozen-API entry, host pointers) on a context created with CKZG_B200_DEVICES=0..%d: "
                            "sharded inside the library, one challenge, one pairing; wall clock of rank 0" % (world - 1)}
                del pin_b, big_b
                ts_all.close()
            except Exception as e:  # noqa: BLE001 -- an extra must not take the headline down
                multi["inlib_multi_device"] = {"error": repr(e)[:300]}
        barrier()

    # ---- N = 1 only: latency / concurrency / coalescing figures ----
    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        # two / three callers sharing the settings (the reference is re-entrant the same way: bindings/go/main_test.go:957-970):
        # the latency-bound tail of one call (transcript, linear combination, pairing) overlaps the throughput kernels of the other
        def _caller(fn, reps):
            for _ in range(reps):
                fn()
        for fn, tag, nthreads in ((verify_dev, "", 2), (verify_dev, "", 3), (verify_host, "e2e_", 2), (verify_host, "e2e_", 3)):
            reps = 6
            th = [threading.Thread(target=_caller, args=(fn, reps)) for _ in range(nthreads)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            extra["verify_blob_kzg_proof_batch_n%d_x%d_concurrent_callers_%sblobs_per_s" % (n, nthreads, tag)] = nthreads * reps * n / (time.perf_counter() - t0)
        # cells only
        for _ in range(1):
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), 0, d_blobs.data_ptr(), m7, ts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), 0, d_blobs.data_ptr(), m7, ts)
        extra["compute_cells_only_batch%d_blobs_per_s" % m7] = m7 * 3 / (time.perf_counter() - t0)
        # verify_cell_kzg_proof_batch at 8 / 64 blobs x 128 cells (one challenge below 2 x 4096 cells)
        for vb in (8, 64):
            idx_arr = (ctypes.c_uint64 * (vb * 128))(*([k for _ in range(vb) for k in range(128)]))
            vc = lambda: mod.verify_cell_kzg_proof_batch_ptr(v_cm.data_ptr(), idx_arr, v_cells.data_ptr(), v_prf.data_ptr(), vb * 128, ts)
            assert vc()
            t0 = time.perf_counter()
            for _ in range(3):
                vc()
            extra["verify_cell_kzg_proof_batch_%dx128_blobs_per_s" % vb] = vb * 3 / (time.perf_counter() - t0)
        mod.profile_enable(ts, 2)
        for _ in range(3):
            vc()
        pv = mod.profile_dump(ts)
        mod.profile_enable(ts, 0)
        extra["verify_cell_64x128_kernels_ms"] = {k: round(v[0] / max(1, v[1]), 3) for k, v in pv["kernels"].items() if k not in ("begin",)}
        # per-blob frozen API from concurrent native host threads (bindings/go/main_test.go:957-970 shape), with and
        # without the call-coalescing front end
        for on in (True, False):
            mod.coalesce_enable(ts, on)
            tag = "coalesced" if on else "uncoalesced"
            mod.bench_per_blob_callers(ts, 1, 8, 1, host_blobs.data_ptr(), 64)
            for nthreads in (8, 64):
                extra["compute_cells_and_kzg_proofs_per_blob_api_%d_threads_%s_blobs_per_s" % (nthreads, tag)] = mod.bench_per_blob_callers(ts, 1, nthreads, 6, host_blobs.data_ptr(), 64)
                extra["blob_to_kzg_commitment_per_blob_api_%d_threads_%s_blobs_per_s" % (nthreads, tag)] = mod.bench_per_blob_callers(ts, 0, nthreads, 24, host_blobs.data_ptr(), 64)
        mod.coalesce_enable(ts, True)
        extra["coalesce_stats(requests,batches,largest)"] = mod.coalesce_stats(ts)
        one = bytes(host_blobs[:BLOB].numpy().tobytes())
        t0 = time.perf_counter()
        for _ in range(3):
            mod.compute_cells_and_kzg_proofs(one, ts)
        extra["compute_cells_and_kzg_proofs_single_call_ms"] = 1000 * (time.perf_counter() - t0) / 3
        for _ in range(2):
            mod.blob_to_kzg_commitment(one, ts)
        t0 = time.perf_counter()
        for _ in range(5):
            mod.blob_to_kzg_commitment(one, ts)
        extra["blob_to_kzg_commitment_single_call_ms"] = 1000 * (time.perf_counter() - t0) / 5

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single_core(bytes(host_blobs.numpy().tobytes()), bytes(host_cms.numpy().tobytes()), bytes(host_prs.numpy().tobytes()), n)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "blobs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms, "engine_ms_per_call": engine_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": workload_config(n, world, args.sharded and world > 1),
            "timing": "CUDA events on the legacy default stream around the K steps (engine call + all-reduce(MIN) + host read of the verdict each step), max over ranks",
            "e2e": {"value": e2e_value, "unit": "blobs/s", "h2d_bytes_per_step": n * (BLOB + 96), "d2h_bytes_per_step": 8 + 64 * n, "ms_per_step": e_ms, "memory": "pinned host", "stages_ms": e2e_stages},
            "e2e_pageable": e2e_pageable, "h2d": h2d,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "commitment": commitment, "configs": configs, "multi_gpu": multi,
            # stage boundaries inside five PROFILED calls right after the timed ones (events on the call's stream, stages
            # overlapping as in production; profiled calls use streams of their own, so their engine_ms_per_call carries
            # the stream creation the timed calls do not pay)
            "stages_ms": {k: round(v[0] / max(1, prof1["calls"]), 4) for k, v in (prof1 or {}).get("kernels", {}).items() if k.startswith("stage:") or k == "end"},
            "extra": extra,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blobs", type=int, default=4096, help="blobs per GPU per step")
    ap.add_argument("--no-extra", action="store_true", help="headline + commitment only")
    ap.add_argument("--sharded", action="store_true", help="N > 1: report the single-challenge sharded global batch as the headline value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
