"""-m gpu: every device primitive against the Python oracle (PTX Montgomery arithmetic, G1 formulas,
decompression / subgroup check), through the engine's self-test entry points."""
import ctypes as C
import random

import pytest

from oracle import bls12_381 as B
from oracle.bls12_381 import P, R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from gpu_common import product_lib_path

    return C.CDLL(product_lib_path())


def pack(vals, n):
    arr = (C.c_uint32 * (n * len(vals)))()
    for i, v in enumerate(vals):
        for j in range(n):
            arr[i * n + j] = (v >> (32 * j)) & 0xFFFFFFFF
    return arr


def unpack(arr, n, count):
    return [sum(int(arr[i * n + j]) << (32 * j) for j in range(n)) for i in range(count)]


def test_fp_field_ops(lib):
    rnd = random.Random(11)
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 380, (1 << 381) % P, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]
    a = edge + [rnd.randrange(P) for _ in range(2000)]
    b = [rnd.choice(edge) for _ in edge] + [rnd.randrange(P) for _ in range(2000)]
    # all edge x edge pairs too
    for x in edge:
        for y in edge:
            a.append(x), b.append(y)
    n = len(a)
    out = (C.c_uint32 * (12 * n))()
    for op, f in [(0, lambda x, y: x * y % P), (1, lambda x, y: (x + y) % P), (2, lambda x, y: (x - y) % P), (6, lambda x, y: x * x % P)]:
        assert lib.ckzg_b200_selftest_field(op, out, pack(a, 12), pack(b, 12), C.c_uint64(n)) == 0
        got = unpack(out, 12, n)
        want = [f(x, y) for x, y in zip(a, b)]
        assert got == want, "Fp op %d" % op
    # binary-Euclid inversion on every operand, plus powers of two (longest shift-only runs)
    a = a + [1 << i for i in range(381)]
    m = len(a)
    out = (C.c_uint32 * (12 * m))()
    assert lib.ckzg_b200_selftest_field(3, out, pack(a, 12), pack(a, 12), C.c_uint64(m)) == 0
    assert unpack(out, 12, m) == [pow(x, P - 2, P) if x else 0 for x in a]


def test_fr_field_ops(lib):
    rnd = random.Random(12)
    edge = [0, 1, 2, R - 1, R - 2, (R - 1) // 2, 1 << 254, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]
    a = edge + [rnd.randrange(R) for _ in range(2000)]
    b = [rnd.choice(edge) for _ in edge] + [rnd.randrange(R) for _ in range(2000)]
    for x in edge:
        for y in edge:
            a.append(x), b.append(y)
    n = len(a)
    out = (C.c_uint32 * (8 * n))()
    assert lib.ckzg_b200_selftest_field(4, out, pack(a, 8), pack(b, 8), C.c_uint64(n)) == 0
    assert unpack(out, 8, n) == [x * y % R for x, y in zip(a, b)]
    a = a + [1 << i for i in range(255)]
    m = len(a)
    out = (C.c_uint32 * (8 * m))()
    assert lib.ckzg_b200_selftest_field(5, out, pack(a, 8), pack(a, 8), C.c_uint64(m)) == 0
    assert unpack(out, 8, m) == [pow(x, R - 2, R) if x else 0 for x in a]


def test_g1_scalar_mul_and_add(lib):
    rnd = random.Random(13)
    n = 48
    pts = [B.g1_mul(B.G1_GEN_J, rnd.randrange(1, R)) for _ in range(n)]
    qs = [B.g1_mul(B.G1_GEN_J, rnd.randrange(1, R)) for _ in range(n)]
    ks = [0, 1, 2, R - 1, R, (1 << 256) - 1] + [rnd.randrange(R) for _ in range(n - 6)]
    qs[0] = B.G1_INF
    pts[7] = B.G1_INF
    qs[9] = B.g1_neg(B.g1_mul(pts[9], ks[9]))  # sum cancels -> infinity
    qs[10] = B.g1_mul(pts[10], ks[10])  # equal points -> doubling branch of madd
    p48 = b"".join(B.g1_compress(p) for p in pts)
    q48 = b"".join(B.g1_compress(q) for q in qs)
    out = C.create_string_buffer(48 * n)
    ok = (C.c_int * n)()
    assert lib.ckzg_b200_selftest_g1(0, out, ok, p48, pack(ks, 8), q48, C.c_uint64(n)) == 0
    for i in range(n):
        assert ok[i] == 1
        want = B.g1_compress(B.g1_add(B.g1_mul(pts[i], ks[i]), qs[i]))
        assert out.raw[48 * i : 48 * i + 48] == want, i


def test_g1_validate_edge_cases(lib):
    """validate_kzg_g1 edge cases (src/test/tests.c:536-745) on the device decompressor."""
    rnd = random.Random(14)
    cases = []  # (bytes, ok_validate, ok_uncompress)
    good = B.g1_compress(B.g1_mul(B.G1_GEN_J, rnd.randrange(1, R)))
    inf = B.g1_compress(B.G1_INF)
    cases += [(good, 1, 1), (inf, 1, 1), (B.g1_compress(B.G1_GEN_J), 1, 1)]
    b = bytearray(good); b[0] &= 0x7F; cases.append((bytes(b), 0, 0))
    b = bytearray(inf); b[0] |= 0x20; cases.append((bytes(b), 0, 0))
    b = bytearray(inf); b[47] = 1; cases.append((bytes(b), 0, 0))
    b = bytearray(inf); b[0] = 0x40; cases.append((bytes(b), 0, 0))
    b = bytearray(P.to_bytes(48, "big")); b[0] |= 0x80; cases.append((bytes(b), 0, 0))
    b = bytearray((P + 1).to_bytes(48, "big")); b[0] |= 0x80; cases.append((bytes(b), 0, 0))
    b = bytearray(48); b[0] = 0x80; cases.append((bytes(b), 0, 0))
    x, n_off, n_out = 5, 0, 0
    while n_off < 4 or n_out < 4:
        x += 1
        y = B.fp_sqrt((x**3 + 4) % P)
        b = bytearray(x.to_bytes(48, "big")); b[0] |= 0x80
        if y is None:
            n_off += 1
            cases.append((bytes(b), 0, 0))
        else:
            in_g1 = B.g1_in_subgroup((x, y, 1))
            n_out += not in_g1
            cases.append((bytes(b), int(in_g1), 1))
    for _ in range(20):
        p = B.g1_to_affine(B.g1_mul(B.G1_GEN_J, rnd.randrange(1, R)))
        for y in (p[1], P - p[1]):
            cases.append((B.g1_compress((p[0], y, 1)), 1, 1))
    n = len(cases)
    p48 = b"".join(c[0] for c in cases)
    out = C.create_string_buffer(48 * n)
    ok = (C.c_int * n)()
    for op, col in ((1, 1), (2, 2)):
        assert lib.ckzg_b200_selftest_g1(op, out, ok, p48, None, None, C.c_uint64(n)) == 0
        for i, c in enumerate(cases):
            assert ok[i] == c[col], (op, i, c[0].hex())
            if c[col]:
                assert out.raw[48 * i : 48 * i + 48] == c[0]
