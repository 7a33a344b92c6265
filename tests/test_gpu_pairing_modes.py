"""-m gpu: the A/B forms of the pairing check give the verdicts of the default form.

`CKZG_B200_PAIRING_MODE` is read once per process (bit 0: final exponentiation on two machines + cooperative inversion,
bit 1: second form of the cyclotomic square, bit 2: packed term codes / tree sums, bit 3: Miller loop on four machines;
default 15, csrc/pairing.cu), so every form runs in a process of its own: a valid blob proof, the proof of another blob,
a two-blob batch with and without a swapped pair, and one verify_kzg_proof vector pair -- true and false cases through
every form.  The default form is pinned against the reference by the golden-vector tests; tests/hostcheck runs the same
forms on the host."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import hashlib, importlib.util, os, sys
root = sys.argv[1]
spec = importlib.util.spec_from_file_location("ckzg_py", os.path.join(root, "c-kzg-4844_b200", "ckzg_py.py"))
ck = importlib.util.module_from_spec(spec); spec.loader.exec_module(ck)
ts = ck.load_trusted_setup()
def blob(seed):
    out = bytearray()
    for i in range(4096):
        out += b"\x00" + hashlib.sha256(b"%d/%d" % (seed, i)).digest()[1:]
    return bytes(out)
b0, b1 = blob(1), blob(2)
c0, c1 = ck.blob_to_kzg_commitment(b0, ts), ck.blob_to_kzg_commitment(b1, ts)
p0, p1 = ck.compute_blob_kzg_proof(b0, c0, ts), ck.compute_blob_kzg_proof(b1, c1, ts)
z = b"\x00" * 31 + b"\x05"
pz, y = ck.compute_kzg_proof(b0, z, ts)
got = [ck.verify_blob_kzg_proof(b0, c0, p0, ts), ck.verify_blob_kzg_proof(b0, c0, p1, ts),
       ck.verify_blob_kzg_proof_batch(b0 + b1, c0 + c1, p0 + p1, ts), ck.verify_blob_kzg_proof_batch(b0 + b1, c0 + c1, p1 + p0, ts),
       ck.verify_kzg_proof(c0, z, y, pz, ts), ck.verify_kzg_proof(c1, z, y, pz, ts)]
print("VERDICTS", "".join("1" if g else "0" for g in got), hashlib.sha256(c0 + c1 + p0 + p1 + pz + y).hexdigest())
'''


def run_mode(mode):
    env = dict(os.environ)
    if mode is None:
        env.pop("CKZG_B200_PAIRING_MODE", None)
    else:
        env["CKZG_B200_PAIRING_MODE"] = str(mode)
    out = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("VERDICTS")][-1].split()
    return line[1], line[2]


def test_every_form_of_the_pairing_check_gives_the_same_verdicts():
    want, digest = run_mode(None)
    assert want == "101010"
    for mode in (0, 1, 3, 7, 14):
        got, d = run_mode(mode)
        assert (got, d) == (want, digest), "CKZG_B200_PAIRING_MODE=%d: %s" % (mode, got)
