#!/usr/bin/env python3
"""Host -> device upload of one step's blobs (512 MiB) from PAGEABLE memory: the staged path of Call::upload for several
thread counts / slot sizes (one subprocess per setting: both are read once per process), cudaHostRegister + direct DMA, and
the pinned-memory DMA beside them.  Usage under gpurun: python tools/upload_probe.py"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BYTES = 512 << 20


def one():
    import numpy as np
    import torch

    import __graft_entry__ as entry

    mod = entry.load_package()
    ts = mod.load_trusted_setup()
    lib = mod.lib()
    ms = C.c_double(0)
    src = np.random.default_rng(1).integers(0, 256, size=BYTES, dtype=np.uint8)
    out = {}
    mode = int(os.environ.get("PROBE_MODE", "0"))
    if mode == 2:
        pin = torch.from_numpy(src).pin_memory()
        assert lib.ckzg_b200_debug_upload(ts.engine, C.c_void_p(pin.data_ptr()), C.c_uint64(BYTES), 5, 0, C.byref(ms)) == 0
    else:
        assert lib.ckzg_b200_debug_upload(ts.engine, C.c_void_p(src.ctypes.data), C.c_uint64(BYTES), 5, mode, C.byref(ms)) == 0
    print(json.dumps({"mode": ["staged", "register", "pinned"][mode], "threads": os.environ.get("CKZG_B200_HOST_THREADS"), "slot_mb": os.environ.get("CKZG_B200_STAGE_SLOT_MB"),
                      "ms": round(ms.value, 3), "GBps": round(BYTES / ms.value / 1e6, 2)}))


if __name__ == "__main__":
    if os.environ.get("PROBE_CHILD"):
        one()
        sys.exit(0)
    print("host cores:", os.cpu_count())
    runs = [{"PROBE_MODE": "2"}, {"PROBE_MODE": "1"}]
    for th in (2, 4, 8, 12, 16):
        for slot in (2, 8, 32):
            runs.append({"PROBE_MODE": "0", "CKZG_B200_HOST_THREADS": str(th), "CKZG_B200_STAGE_SLOT_MB": str(slot)})
    for env in runs:
        r = subprocess.run([sys.executable, __file__], env=dict(os.environ, PROBE_CHILD="1", **env), capture_output=True, text=True, timeout=300)
        print((r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1])
