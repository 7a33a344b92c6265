"""A pure-C consumer of the frozen API (examples/c_consumer.c) compiled with gcc against include/ckzg.h and
linked with libckzg_b200.so: the drop-in boundary seen from the reference's own host language.  Without a GPU
the program must report C_KZG_ERROR from load_trusted_setup_file (no CPU fallback); on a B200 it must run every
API entry and agree with itself."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "c-kzg-4844_b200")
EXE = os.path.join(ROOT, "examples", "_build", "c_consumer")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_consumer.c"), "-L", LIBDIR, "-lckzg_b200",
           "-Wl,-rpath," + LIBDIR, "-o", EXE]
    subprocess.check_call(cmd)
    return EXE


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def test_c_consumer_builds_and_fails_loudly_without_a_device():
    if not os.path.exists(os.path.join(LIBDIR, "libckzg_b200.so")):
        pytest.skip("library not built")
    exe = build()
    if has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, os.path.join(LIBDIR, "data", "trusted_setup.txt")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "C_KZG_ERROR" in r.stderr, (r.returncode, r.stdout, r.stderr)


def test_pkg_config_file_matches_the_tree():
    pc = os.path.join(LIBDIR, "ckzg_b200.pc")
    if not os.path.exists(pc):
        pytest.skip("library not built")
    text = open(pc).read()
    assert "-lckzg_b200" in text and os.path.join(ROOT, "include") in text
    if shutil.which("pkg-config"):
        out = subprocess.check_output(["pkg-config", "--cflags", "--libs", pc], text=True)
        assert "-lckzg_b200" in out


@pytest.mark.gpu
def test_c_consumer_runs_every_api_entry():
    exe = build()
    r = subprocess.run([exe, os.path.join(LIBDIR, "data", "trusted_setup.txt")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "c_consumer: ok" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_consumer_batch_phase_on_a_multi_device_context():
    """The same C program, unchanged, with the library told to span two devices (CKZG_B200_DEVICES; on a one-GPU box
    two replicas share device 0 -- same code path): the 640-blob verify_blob_kzg_proof_batch and the 64 x 128-cell
    verify_cell_kzg_proof_batch are sharded inside the library and give the same verdicts, incl. negative controls."""
    import torch

    exe = build()
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    env = dict(os.environ, CKZG_B200_DEVICES=devs, CKZG_B200_DEBUG="1")
    r = subprocess.run([exe, os.path.join(LIBDIR, "data", "trusted_setup.txt"), "640"], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "c_consumer: ok" in r.stdout and "batch phase n=640 (8192 cells) ok" in r.stdout, (r.returncode, r.stdout, r.stderr[-2000:])
    assert "context spans 2 devices" in r.stderr, r.stderr[-2000:]
    assert "verify_blob_kzg_proof_batch: 640 blobs over 2 devices" in r.stderr, r.stderr[-2000:]
