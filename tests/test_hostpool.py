"""CPU tier: the host worker pool of the C-ABI layer (c-kzg-4844_b200/csrc/hostpool.h) under nested and concurrent
use -- it carries the parallel memcpy of pageable caller buffers into pinned staging and nothing else may stall on it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hostpool_every_index_once_nested_and_concurrent(tmp_path):
    exe = str(tmp_path / "hostpool_test")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(
        ["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "c-kzg-4844_b200", "csrc"), "-I", os.path.join(cuda, "include"),
         os.path.join(ROOT, "tests", "hostcheck", "hostpool_test.cpp"), "-o", exe, "-lpthread", "-L", os.path.join(cuda, "lib64"), "-lcudart"]
    )
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and "ok bad=0" in out.stdout, out.stdout + out.stderr
