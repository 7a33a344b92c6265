"""The one host-side hash (c-kzg-4844_b200/src/host_sha256.c: SHA-NI, fully unrolled; portable fallback) against
hashlib: every length around the block and padding boundaries, and the update patterns of the two transcripts
(160-byte records, eip4844.c:648-660; 16 + 2048 + 48-byte records, eip7594.c:405-474).  CPU only -- the source is
compiled on its own, the product library is not needed."""
import ctypes as C
import hashlib
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.path.dirname(HERE), "c-kzg-4844_b200", "src")
OUT = os.path.join(HERE, "hostcheck", "_build", "libhostsha.so")


@pytest.fixture(scope="module", params=["native", "portable"])
def sha(request):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    out = OUT.replace(".so", "_%s.so" % request.param)
    flags = [] if request.param == "native" else ["-DCKZG_HOST_SHA_PORTABLE"]
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-I", SRC] + flags + ["-o", out, os.path.join(SRC, "host_sha256.c")])
    return C.CDLL(out)


def digest(lib, pieces):
    st = C.create_string_buffer(256)
    lib.ckzg_host_sha256_init(st)
    for p in pieces:
        lib.ckzg_host_sha256_update(st, p, C.c_size_t(len(p)))
    out = C.create_string_buffer(32)
    lib.ckzg_host_sha256_final(st, out)
    return out.raw


def test_lengths_around_block_boundaries(sha):
    rnd = random.Random(1)
    for n in list(range(0, 200)) + [255, 256, 257, 1000, 4095, 4096, 4097, 131152]:
        msg = bytes(rnd.randrange(256) for _ in range(n))
        assert digest(sha, [msg]) == hashlib.sha256(msg).digest(), n


def test_known_answers(sha):
    assert digest(sha, [b"abc"]).hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert digest(sha, [b""]).hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"


def test_transcript_update_patterns(sha):
    rnd = random.Random(2)
    # blob-batch transcript: 32-byte head + n records of 160 bytes
    pieces = [bytes(rnd.randrange(256) for _ in range(32))] + [bytes(rnd.randrange(256) for _ in range(160)) for _ in range(37)]
    assert digest(sha, pieces) == hashlib.sha256(b"".join(pieces)).digest()
    # cell-batch transcript: 48-byte head, u commitments, then per cell 16 + 2048 + 48 bytes
    pieces = [bytes(48), bytes(rnd.randrange(256) for _ in range(48 * 3))]
    for _ in range(21):
        pieces += [bytes(rnd.randrange(256) for _ in range(16)), bytes(rnd.randrange(256) for _ in range(2048)), bytes(rnd.randrange(256) for _ in range(48))]
    assert digest(sha, pieces) == hashlib.sha256(b"".join(pieces)).digest()
    # arbitrary splits
    msg = bytes(rnd.randrange(256) for _ in range(5000))
    cuts = sorted(rnd.sample(range(5000), 40))
    parts = [msg[a:b] for a, b in zip([0] + cuts, cuts + [5000])]
    assert digest(sha, parts) == hashlib.sha256(msg).digest()
