#!/usr/bin/env python3
"""bench.py -- blobs/sec of the KZG hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, all host cores

Headline workload: verify_blob_kzg_proof_batch at the north-star batch size (4096 blobs per GPU), the
configuration BASELINE.json's metric and target are quoted on; configs[1] (n=64) and
blob_to_kzg_commitment are measured too and reported under "extra".  One "step" = one pass of the hot
path over one batch of synthetic blobs.

  value ........ blobs/s with blobs, commitments and proofs already resident in HBM (device pointers
                 into the engine's C ABI), CUDA-event time of the call on the stream it launches on
  e2e .......... the same call with HOST pointers (pinned): H2D of every input and D2H of the verdict
                 inside the timed region
  roofline ..... dominant kernel (largest share of the step), timed live by the engine's event trace
  cpu_baseline . the unmodified reference (oracle/_ref) on 1 host core, bounded sample (rank 0, N=1)

Multi-GPU (--gpus N under torchrun): weak scaling, each rank verifies its own 4096-blob batch and the
verdicts are combined with one NCCL all-reduce(MIN) per step inside the timed region
(c-kzg-4844_b200/parallel.py; the single-challenge sharded mode is `--sharded`).
Inputs are larger than L2 (512 MiB of blobs per step vs 126 MB), so no explicit flush is needed.
"""
import argparse
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
    "pack_tuples": 320,
    "r_challenge": 160,
    "pairing_check": 0,
    "msm_sort": 131072 + 393216 + 4100,
    "msm_accumulate": 393216 + 4100 + 196608,  # digit lists + bucket offsets in, 1024 XYZZ buckets out (table gathers are L2 hits)
    "msm_reduce": 196608 + 192,
    "g1_compress": 192 + 48,
}
# measured DRAM traffic per blob (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
# capture, profiles/r01_summary.md "r01k"/"r01o"), for the kernels that were captured
MEASURED_TRAFFIC_PER_BLOB = {"msm_accumulate": 596_000, "hash+validate": 135_400, "rlc_points": 900}  # hash+validate: r01t capture (539 MB + 15 MB per 4096 blobs)

# algorithmic 32x32->64 multiply-accumulates per blob (SURVEY.md section 8(d) convention: Fp mul = 300, Fr mul = 136)
ALGO_MAC_PER_BLOB = {
    "evaluate": (4096 + 2 * 4095) * 136,   # tree evaluation: one product per leaf, two per node
    "g1_validate": 2 * 2200 * 300,
    "hash+validate": 2 * 2200 * 300,
    "rlc_points": 3 * 3600 * 300,
    "vmsm_accumulate": 3 * 32 * 14 * 300,  # ~32 non-zero signed bytes per scalar, three scalars per blob, 14 products per XYZZ addition
    "msm_accumulate": 98304 * 10 * 300,
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
def _ref_worker(args):
    wid, steps, warmup, n_distinct, tile = args
    import numpy as np

    from oracle import ref_lib

    k = ref_lib.CKZG()
    blobs = synth_blobs(n_distinct, 4844 + 1000 + wid).tobytes()
    bl = [blobs[BLOB * i : BLOB * (i + 1)] for i in range(n_distinct)]
    cms = [k.blob_to_kzg_commitment(b) for b in bl]
    prs = [k.compute_blob_kzg_proof(b, c) for b, c in zip(bl, cms)]
    B, Cc, P = blobs * tile, b"".join(cms) * tile, b"".join(prs) * tile
    n = n_distinct * tile
    for _ in range(warmup):
        assert k.verify_blob_kzg_proof_batch(B, Cc, P)
    t0 = time.perf_counter()
    for _ in range(steps):
        assert k.verify_blob_kzg_proof_batch(B, Cc, P)
    return n * steps, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import ref_lib

    if not os.path.exists(ref_lib.REF_SO):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libckzg_ref.so missing (build with oracle/build_ref.sh)"}))
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n_distinct, tile = 32, 16  # n = 512 per call per worker
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, [(w, args.steps, args.warmup, n_distinct, tile) for w in range(cores)])
    total = sum(r[0] for r in res)
    tmax = max(r[1] for r in res)
    value = total / tmax
    sample = "each of %d processes: verify_blob_kzg_proof_batch n=%d (%d distinct synthetic blobs tiled x%d), %d calls" % (cores, n_distinct * tile, n_distinct, tile, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "blobs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * tmax / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (381/255-bit Montgomery integers)",
        "data": "synthetic", "config": {"workload": "verify_blob_kzg_proof_batch n=4096 per GPU (reference arm: bounded sample, see cpu_baseline.sample)"},
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
    mod = entry.load_package()
    par = __import__("importlib").import_module("ckzg_b200.parallel")
    ts = mod.load_trusted_setup()
    n = args.blobs

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

    # --sharded: ONE global batch of world*n blobs with a single Fiat-Shamir challenge (exact reference
    # semantics for the concatenation): per-blob stage on each rank, all-gather of z||y (64 B/blob),
    # partial linear combinations, all-gather of 2 compressed points per rank, one pairing.
    all_cms = all_prs = None
    if args.sharded:
        if world > 1:
            gc = [torch.empty_like(d_cms) for _ in range(world)]
            gp = [torch.empty_like(d_prs) for _ in range(world)]
            dist.all_gather(gc, d_cms)
            dist.all_gather(gp, d_prs)
            all_cms = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gc)
            all_prs = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gp)
        else:
            all_cms, all_prs = bytes(host_cms.numpy().tobytes()), bytes(host_prs.numpy().tobytes())

    def verify_sharded():
        if world == 1:
            zy = mod.verify_stage1(d_blobs.data_ptr(), d_cms.data_ptr(), d_prs.data_ptr(), n, ts)
            tuples = b"".join(all_cms[48 * i : 48 * i + 48] + zy[64 * i : 64 * i + 64] + all_prs[48 * i : 48 * i + 48] for i in range(n))
            return mod.verify_finish(mod.verify_stage2(tuples, n, 0, n, ts), 1, ts)
        return par.verify_batch_sharded(
            lambda: mod.verify_stage1(d_blobs.data_ptr(), d_cms.data_ptr(), d_prs.data_ptr(), n, ts),
            lambda tuples, nt, first, nl: mod.verify_stage2(tuples, nt, first, nl, ts),
            lambda parts, nr: mod.verify_finish(parts, nr, ts),
            all_cms, all_prs, world * n, dev,
        )

    def step(fn):
        if args.sharded and fn is verify_dev:
            ok = verify_sharded()
        else:
            ok = par.verify_batch_replicas(fn, dev)
        assert ok, "verification of valid synthetic batch failed"

    # negative control: two swapped proofs must be rejected
    bad = d_prs.clone()
    bad[0:48], bad[48:96] = d_prs[48:96].clone(), d_prs[0:48].clone()
    assert not mod.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_cms.data_ptr(), bad.data_ptr(), n, ts), "negative control accepted"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile):
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(warmup):
            step(fn)
        barrier()
        if profile:
            mod.profile_enable(ts, profile)
        l0 = ts.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            step(fn)
        barrier()
        t1 = time.perf_counter()
        wall = t1 - t0
        clocks = sampler.stop(t0, t1)
        prof = mod.profile_dump(ts) if profile else None
        if profile:
            mod.profile_enable(ts, 0)
        launches = ts.launch_count() - l0
        dev_ms = (prof["call_ms"] / steps) if prof else None
        # max over ranks of both clocks
        t = torch.tensor([wall, dev_ms if dev_ms is not None else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), prof, launches, clocks

    # timed run: whole-call CUDA events only (level 1), stages free to overlap
    wall, dev_ms, prof1, launches, clocks = timed(verify_dev, args.steps, args.warmup, 1)
    # kernel breakdown: separate short run with per-kernel events (level 2, stages serialised)
    _, _, prof, _, _ = timed(verify_dev, 2, 1, 2)
    prof_steps = 2
    # device-resident number: CUDA events on the launching stream (engine trace), max over ranks
    ms_per_step = dev_ms
    value = world * n / (ms_per_step / 1000.0)
    e_wall, _, _, _, _ = timed(verify_host, args.steps, max(1, args.warmup - 1), 0)
    e2e_value = world * n / (e_wall / args.steps)

    # ---- roofline of the dominant kernel ----
    peak, peak_src = load_peaks()
    kern = {k: v for k, v in prof["kernels"].items() if k not in ("begin", "end")}
    total_ms = sum(v[0] for v in kern.values()) or 1.0
    # The roofline kernel is the largest of the kernels that stream the per-blob bytes.  At n=4096 the
    # per-call tail (one pairing check, the transcript hash, the tree sums) is comparable in time but
    # moves no per-blob data, so an HBM figure for it would be meaningless; it is named separately.
    per_call = ("pairing_check", "r_from_digest", "g1_sum", "transcript(d2h,host_sha)", "rlc_scalars", "vmsm_sort", "vmsm_accumulate", "vmsm_reduce")
    # "stream the per-blob bytes" = read the blob itself (131,168 B per blob for this path)
    streaming = [k for k in kern if k not in per_call and ALGO_BYTES_PER_BLOB.get(k, 0) >= 100_000] or list(kern)
    dom = max(streaming, key=lambda k: kern[k][0])
    time_dom = max(kern, key=lambda k: kern[k][0])
    dom_ms_per_launch = kern[dom][0] / kern[dom][1]
    launches_per_step = kern[dom][1] / prof_steps
    units_per_launch = n / launches_per_step
    algo_bytes = ALGO_BYTES_PER_BLOB.get(dom, 0) * units_per_launch
    achieved = algo_bytes / (dom_ms_per_launch * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (MEASURED_TRAFFIC_PER_BLOB[dom] * units_per_launch) if dom in MEASURED_TRAFFIC_PER_BLOB else None,
        "peak_source": peak_src, "share_of_step": kern[dom][0] / total_ms, "ms_per_launch": dom_ms_per_launch,
        "algorithmic_bytes_per_blob": ALGO_BYTES_PER_BLOB.get(dom, 0),
        "largest_kernel_by_time": {"kernel": time_dom, "share_of_step": kern[time_dom][0] / total_ms,
                                   "per_call_fixed_work": time_dom in per_call},
        "note": "integer-pipe bound path (SURVEY 8d): HBM fraction is reported for completeness, see int_pipe",
    }
    sm_mhz = clocks.get("sm_mhz") or 1965
    peak_mac = 148 * 64 * sm_mhz * 1e6
    int_pipe = {}
    for k, mac in ALGO_MAC_PER_BLOB.items():
        if k in kern and kern[k][0] > 0:
            t_s = kern[k][0] / prof_steps * 1e-3
            int_pipe[k] = {"mac_per_s": mac * n / t_s, "frac_of_peak": mac * n / t_s / peak_mac,
                           # the multiplier microbenchmark (tools/gpu_probe.py mulbench, profiles/) tops out at
                           # 9.34e12 MAC/s at 1965 MHz = 32 wide MAC/clk/SM: the attainable ceiling for this code
                           "frac_of_measured_mul_peak": mac * n / t_s / (9.34e12 * sm_mhz / 1965.0)}
    shares = {k: round(v[0] / total_ms, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
    roofline["per_kernel"] = {
        k: {"ms_per_step": round(v[0] / prof_steps, 4), "algorithmic_GBps": round(ALGO_BYTES_PER_BLOB.get(k, 0) * n / (v[0] / prof_steps * 1e-3) / 1e9, 3) if v[0] > 0 else None}
        for k, v in kern.items()
    }

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:  # the other configs / APIs: single-GPU runs only
        # configs[1]: n = 64 in one call (latency-bound case), and commitments
        def v64():
            return mod.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_cms.data_ptr(), d_prs.data_ptr(), 64, ts)
        for _ in range(3):
            assert v64()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            v64()
        extra["verify_blob_kzg_proof_batch_n64_blobs_per_s"] = 64 * 5 / (time.perf_counter() - t0)
        # two callers sharing the settings (the reference is re-entrant the same way: bindings/go/main_test.go:957-970):
        # the latency-bound tail of one call (challenge transcript, RLC, pairing) overlaps the throughput
        # kernels of the other
        import threading

        # Throughput with several host threads calling into the SAME settings object (each call has its own
        # stream and scratch): the per-call tail -- transcript hash, linear combination, pairing -- is
        # latency-bound and leaves most SMs idle, so a second caller's per-blob stage (and, through host
        # pointers, its upload) runs underneath it.  The headline numbers above stay single-caller.
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
        m = min(n, 1024)
        out = torch.empty(48 * m, dtype=torch.uint8, device=dev)
        for _ in range(2):
            mod.blob_to_kzg_commitment_device(out.data_ptr(), d_blobs.data_ptr(), m, ts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            mod.blob_to_kzg_commitment_device(out.data_ptr(), d_blobs.data_ptr(), m, ts)
        extra["blob_to_kzg_commitment_batch%d_blobs_per_s" % m] = m * 3 / (time.perf_counter() - t0)
        # EIP-7594 (north_star target 1): BASELINE configs[2] shape = 256 blobs, cells + FK20 proofs
        m7 = min(n, 256)
        d_cells = torch.empty(m7 * 2 * BLOB, dtype=torch.uint8, device=dev)
        d_cprf = torch.empty(m7 * 128 * 48, dtype=torch.uint8, device=dev)
        mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), d_cprf.data_ptr(), d_blobs.data_ptr(), m7, ts)
        torch.cuda.synchronize()
        mod.profile_enable(ts, 2)
        t0 = time.perf_counter()
        for _ in range(3):
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), d_cprf.data_ptr(), d_blobs.data_ptr(), m7, ts)
        dt = (time.perf_counter() - t0) / 3
        p7 = mod.profile_dump(ts)
        mod.profile_enable(ts, 0)
        extra["compute_cells_and_kzg_proofs_batch%d_blobs_per_s" % m7] = m7 / dt
        extra["compute_cells_and_kzg_proofs_kernels_ms"] = {k: round(v[0] / 3, 3) for k, v in p7["kernels"].items() if k not in ("begin", "end")}
        h_cells = torch.empty(m7 * 2 * BLOB, dtype=torch.uint8).pin_memory()
        h_cprf = torch.empty(m7 * 128 * 48, dtype=torch.uint8).pin_memory()
        mod.compute_cells_and_kzg_proofs_host(h_cells.data_ptr(), h_cprf.data_ptr(), host_blobs.data_ptr(), m7, ts)
        t0 = time.perf_counter()
        for _ in range(3):
            mod.compute_cells_and_kzg_proofs_host(h_cells.data_ptr(), h_cprf.data_ptr(), host_blobs.data_ptr(), m7, ts)
        extra["compute_cells_and_kzg_proofs_batch%d_e2e_blobs_per_s" % m7] = m7 * 3 / (time.perf_counter() - t0)
        for _ in range(1):
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), 0, d_blobs.data_ptr(), m7, ts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            mod.compute_cells_and_kzg_proofs_device(d_cells.data_ptr(), 0, d_blobs.data_ptr(), m7, ts)
        extra["compute_cells_only_batch%d_blobs_per_s" % m7] = m7 * 3 / (time.perf_counter() - t0)
        # configs[3] shape: recover from the even-indexed cells (50 % missing)
        idx = list(range(0, 128, 2)) * m7
        cells_view = d_cells.view(m7, 128, 2048)
        given = cells_view[:, 0::2, :].contiguous()
        rec_c = torch.empty(m7 * 2 * BLOB, dtype=torch.uint8, device=dev)
        rec_p = torch.empty(m7 * 128 * 48, dtype=torch.uint8, device=dev)
        mod.recover_cells_and_kzg_proofs_device(rec_c.data_ptr(), rec_p.data_ptr(), idx, given.data_ptr(), 64, m7, ts)
        torch.cuda.synchronize()
        assert torch.equal(rec_c, d_cells) and torch.equal(rec_p, d_cprf), "recover != compute"
        t0 = time.perf_counter()
        for _ in range(3):
            mod.recover_cells_and_kzg_proofs_device(rec_c.data_ptr(), rec_p.data_ptr(), idx, given.data_ptr(), 64, m7, ts)
        extra["recover_cells_and_kzg_proofs_half_missing_batch%d_blobs_per_s" % m7] = m7 * 3 / (time.perf_counter() - t0)
        # verify_cell_kzg_proof_batch: 8 blobs x 128 cells through the frozen API (host bytes)
        nb = 8
        hc = h_cells.numpy().tobytes()
        hp = h_cprf.numpy().tobytes()
        cm_host = bytes(host_cms.numpy().tobytes())
        vc_cm = [cm_host[48 * b : 48 * b + 48] for b in range(nb) for _ in range(128)]
        vc_idx = [k for _ in range(nb) for k in range(128)]
        vc_cells = [hc[2048 * i : 2048 * (i + 1)] for i in range(nb * 128)]
        vc_prf = [hp[48 * i : 48 * (i + 1)] for i in range(nb * 128)]
        assert mod.verify_cell_kzg_proof_batch(vc_cm, vc_idx, vc_cells, vc_prf, ts)
        mod.profile_enable(ts, 2)
        t0 = time.perf_counter()
        for _ in range(3):
            mod.verify_cell_kzg_proof_batch(vc_cm, vc_idx, vc_cells, vc_prf, ts)
        extra["verify_cell_kzg_proof_batch_8x128_blobs_per_s"] = nb * 3 / (time.perf_counter() - t0)
        pv = mod.profile_dump(ts)
        mod.profile_enable(ts, 0)
        extra["verify_cell_kernels_ms"] = {k: round(v[0] / 3, 3) for k, v in pv["kernels"].items() if k not in ("begin",)}
        extra["verify_cell_device_ms_per_call"] = pv["call_ms"] / 3
        # larger batches through the C ABI with host pointers (row-major by blob, bindings/go/main_test.go:1016-1028):
        # the whole input enters ONE serial transcript hash on the host (eip7594.c:405-474), 2112 B per cell
        import ctypes

        h_cm_rows = host_cms.view(-1, 48)[:m7].repeat_interleave(128, dim=0).contiguous().pin_memory()
        for nb in (64, m7):
            if nb > m7:
                continue
            idx_arr = (ctypes.c_uint64 * (nb * 128))(*([k for _ in range(nb) for k in range(128)]))
            def vc_call():
                return mod.verify_cell_kzg_proof_batch_ptr(h_cm_rows.data_ptr(), idx_arr, h_cells.data_ptr(), h_cprf.data_ptr(), nb * 128, ts)
            assert vc_call()
            mod.profile_enable(ts, 2)
            t0 = time.perf_counter()
            for _ in range(3):
                vc_call()
            extra["verify_cell_kzg_proof_batch_%dx128_blobs_per_s" % nb] = nb * 3 / (time.perf_counter() - t0)
            pv = mod.profile_dump(ts)
            mod.profile_enable(ts, 0)
            extra["verify_cell_%dx128_kernels_ms" % nb] = {k: round(v[0] / 3, 3) for k, v in pv["kernels"].items() if k not in ("begin",)}
        # per-blob frozen API from concurrent host threads (bindings/go/main_test.go:957-970 shape), with and
        # without the call-coalescing front end
        # (native host threads: Python threads would serialise on the interpreter lock around each call)
        blob_bytes = [bytes(host_blobs[:BLOB].numpy().tobytes())]
        for on in (True, False):
            mod.coalesce_enable(ts, on)
            tag = "coalesced" if on else "uncoalesced"
            mod.bench_per_blob_callers(ts, 1, 8, 1, host_blobs.data_ptr(), 64)
            for nthreads in (8, 64):
                extra["compute_cells_and_kzg_proofs_per_blob_api_%d_threads_%s_blobs_per_s" % (nthreads, tag)] = mod.bench_per_blob_callers(ts, 1, nthreads, 6, host_blobs.data_ptr(), 64)
                extra["blob_to_kzg_commitment_per_blob_api_%d_threads_%s_blobs_per_s" % (nthreads, tag)] = mod.bench_per_blob_callers(ts, 0, nthreads, 24, host_blobs.data_ptr(), 64)
        mod.coalesce_enable(ts, True)
        extra["coalesce_stats(requests,batches,largest)"] = mod.coalesce_stats(ts)
        t0 = time.perf_counter()
        for _ in range(3):
            mod.compute_cells_and_kzg_proofs(blob_bytes[0], ts)
        extra["compute_cells_and_kzg_proofs_single_call_ms"] = 1000 * (time.perf_counter() - t0) / 3
        one = bytes(host_blobs[:BLOB].numpy().tobytes())
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
            "ms_per_step": ms_per_step, "wall_ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (381/255-bit Montgomery integers)", "data": "synthetic",
            "config": {
                "workload": "verify_blob_kzg_proof_batch n=%d per GPU (north_star batch; configs[1] n=64 under extra)" % n,
                "blobs_per_gpu": n, "parallelism": ("sharded global batch x%d: all-gather(z||y) + all-gather(partials)" % world) if args.sharded else ("replicas x%d + 1 all-reduce(MIN)" % world if world > 1 else "single GPU"),
                "l2": "inputs (%.0f MiB/step) exceed the 126 MB L2; no explicit flush" % (n * BLOB / 2**20),
            },
            "e2e": {"value": e2e_value, "unit": "blobs/s", "h2d_bytes_per_step": n * (BLOB + 96), "d2h_bytes_per_step": 8, "ms_per_step": 1000.0 * e_wall / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "int_pipe": int_pipe, "kernel_share": shares,
            # stage boundaries inside the TIMED calls (events on the call's stream, stages overlapping as in production)
            "stages_ms": {k: round(v[0] / args.steps, 4) for k, v in (prof1 or {}).get("kernels", {}).items() if k.startswith("stage:") or k == "end"},
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
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--sharded", action="store_true", help="one global batch, single challenge, all-gather exchange (parallel.verify_batch_sharded)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
