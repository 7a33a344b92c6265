"""Build the TEST-ONLY host harness (tests/hostcheck/hostcheck.cpp) with g++."""
import os, subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "c-kzg-4844_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libhostcheck.so")


def build(pairing=True, defines=(), tag=""):
    """`defines`: extra -D switches (A/B forms of the arithmetic that are off in the product build), `tag`: output suffix."""
    out = OUT.replace(".so", tag + ".so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(HERE, "hostcheck.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", CSRC, "-o", out, os.path.join(HERE, "hostcheck.cpp")]
    cmd[1:1] = ["-D" + d for d in defines]
    if pairing and os.path.exists(os.path.join(CSRC, "pairing.cuh")):
        cmd.insert(1, "-DHOSTCHECK_PAIRING")
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build())
