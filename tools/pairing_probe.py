#!/usr/bin/env python3
"""Where the time of one pairing check goes: clock64() marks inside the flow of pairing_check_kernel and timed loops
of every cooperative tower operation (ckzg_b200_debug_pairing_probe, csrc/pairing.cu).  No torch: ctypes on the library.
Usage under gpurun: python tools/pairing_probe.py [reps]"""
import ctypes as C
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "c-kzg-4844_b200")
spec = importlib.util.spec_from_file_location("ckzg_py", os.path.join(PKG, "ckzg_py.py"))
ck = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ck)

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ts = ck.load_trusted_setup()
with open(ck.SETUP_TXT) as f:
    tok = f.read().split()
n1 = int(tok[0])
pts = bytes.fromhex(tok[2] + tok[3])  # two Lagrange-form setup points: valid subgroup elements
ticks = (C.c_longlong * 64)()
ok = C.c_int(-1)
fn = ck.lib().ckzg_b200_debug_pairing_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_char_p, C.c_int]
for rnd in range(2):  # second run: warm instruction cache
    rc = fn(ts.engine, ticks, C.byref(ok), pts, reps)
    assert rc == 0, rc
t = list(ticks)
MHZ = float(os.environ.get("SM_MHZ", "1965"))
us = lambda a, b: (t[b] - t[a]) / MHZ
names = ["inputs (tables, points, 2 x 68 lines)", "Miller loops (two machines side by side)", "F0 * F1", "(tick)", "Fp12 inversion (serial lane)",
         "rest of the easy part (conj, 2 mul, frobenius)", "first pow_x + conj + mul", "rest of the hard part", "final comparison"]
print("verdict", ok.value, "(two unrelated points: 0 expected)   total %.1f us" % us(0, 9))
for i, nm in enumerate(names):
    print("  %-52s %9.1f us" % (nm, us(i, i + 1)))
ops = ["cyclotomic square", "product", "square", "line product", "conjugation", "copy", "frobenius", "inversion"]
for k, nm in enumerate(ops):
    r = min(reps, 2) if k == 7 else reps
    print("  op %-20s %8.2f us each (%d reps)" % (nm, us(16 + 2 * k, 17 + 2 * k) / r, r))
