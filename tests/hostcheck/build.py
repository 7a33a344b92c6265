"""Build the TEST-ONLY host harness (tests/hostcheck/hostcheck.cpp) with g++."""
import os, subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "c-kzg-4844_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libhostcheck.so")


def build(pairing=True):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(HERE, "hostcheck.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", CSRC, "-o", OUT, os.path.join(HERE, "hostcheck.cpp")]
    if pairing and os.path.exists(os.path.join(CSRC, "pairing.cuh")):
        cmd.insert(1, "-DHOSTCHECK_PAIRING")
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build())
