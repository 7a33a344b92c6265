"""Formula-level checks of the engine's header-only arithmetic, host-compiled (portable path of
field.cuh) and compared with the Python oracle.  CPU only; the PTX paths are covered by the -m gpu
tests (tests/test_gpu_units.py)."""
import ctypes as C
import random

import pytest

from oracle import bls12_381 as B
from oracle.bls12_381 import P, R

from hostcheck.build import build


@pytest.fixture(scope="module")
def hc():
    return C.CDLL(build())


def limbs(v, n):
    return (C.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(arr):
    return sum(int(x) << (32 * i) for i, x in enumerate(arr))


EDGE_P = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, (1 << 380), (1 << 381) - 1 - (1 << 381) % 1]
EDGE_R = [0, 1, 2, R - 1, R - 2, (R - 1) // 2, 1 << 254, 0xFFFFFFFF, 1 << 32]


def test_fp_ops(hc):
    rnd = random.Random(1)
    vals = [v % P for v in EDGE_P] + [rnd.randrange(P) for _ in range(60)]
    out, o2, o3 = (C.c_uint32 * 12)(), (C.c_uint32 * 12)(), (C.c_uint32 * 12)()
    for a in vals:
        for b in vals[:12] + [rnd.randrange(P)]:
            hc.hc_fp_mul(out, limbs(a, 12), limbs(b, 12))
            assert val(out) == a * b % P
            hc.hc_fp_addsub(out, o2, o3, limbs(a, 12), limbs(b, 12))
            assert val(out) == (a + b) % P and val(o2) == (a - b) % P and val(o3) == (-a) % P
    # binary-Euclid inversion: all edge values, small values, powers of two, and a long random run
    for a in vals + list(range(1, 40)) + [1 << i for i in range(381)] + [rnd.randrange(P) for _ in range(1500)]:
        a %= P
        hc.hc_fp_inv(out, limbs(a, 12))
        assert val(out) == (pow(a, P - 2, P) if a else 0)


def test_fr_ops(hc):
    rnd = random.Random(2)
    vals = EDGE_R + [rnd.randrange(R) for _ in range(60)]
    out = (C.c_uint32 * 8)()
    for a in vals:
        for b in vals[:10] + [rnd.randrange(R)]:
            hc.hc_fr_mul(out, limbs(a, 8), limbs(b, 8))
            assert val(out) == a * b % R
    for a in vals + list(range(1, 40)) + [1 << i for i in range(255)] + [rnd.randrange(R) for _ in range(1500)]:
        a %= R
        hc.hc_fr_inv(out, limbs(a, 8))
        assert val(out) == (pow(a, R - 2, R) if a else 0)


def test_fr_bytes(hc):
    rnd = random.Random(3)
    ob = C.create_string_buffer(32)
    for v in [0, 1, R - 1, R, R + 1, (1 << 256) - 1, 2 * R, 2 * R + 5] + [rnd.randrange(1 << 256) for _ in range(50)]:
        ok = hc.hc_fr_from_be(ob, v.to_bytes(32, "big"))
        assert bool(ok) == (v < R)
        if v < R:
            assert ob.raw == v.to_bytes(32, "big")
        hc.hc_fr_hash_reduce(ob, v.to_bytes(32, "big"))
        assert ob.raw == (v % R).to_bytes(32, "big")


def rand_g1(rnd):
    return B.g1_mul(B.G1_GEN_J, rnd.randrange(1, R))


def test_g1_mul_add_compress(hc):
    rnd = random.Random(4)
    out = C.create_string_buffer(48)
    inf = B.g1_compress(B.G1_INF)
    for it in range(12):
        p = rand_g1(rnd)
        q = rand_g1(rnd)
        k = [0, 1, 2, R - 1, R, rnd.randrange(R), rnd.randrange(1 << 256)][it % 7]
        pc, qc = B.g1_compress(p), B.g1_compress(q)
        assert hc.hc_g1_mul_add(out, pc, limbs(k, 8), qc) == 1
        assert out.raw == B.g1_compress(B.g1_add(B.g1_mul(p, k), q))
        assert hc.hc_g1_mul_add(out, pc, limbs(k, 8), None) == 1
        assert out.raw == B.g1_compress(B.g1_mul(p, k))
    # infinity inputs
    assert hc.hc_g1_mul_add(out, inf, limbs(5, 8), inf) == 1 and out.raw == inf


def test_g1_madd_special_cases(hc):
    rnd = random.Random(5)
    out = C.create_string_buffer(48)
    p = rand_g1(rnd)
    q = rand_g1(rnd)
    pc, qc, inf = B.g1_compress(p), B.g1_compress(q), B.g1_compress(B.G1_INF)
    cases = [
        (pc, qc, 0, B.g1_add(p, q)),
        (pc, qc, 1, B.g1_sub(p, q)),
        (pc, pc, 0, B.g1_dbl(p)),  # acc == a -> doubling branch
        (pc, pc, 1, B.G1_INF),  # acc == -a -> infinity
        (inf, qc, 0, q),
        (inf, qc, 1, B.g1_neg(q)),
        (pc, inf, 0, p),
    ]
    for a, b, negf, want in cases:
        assert hc.hc_g1_madd(out, a, b, negf) == 1
        assert out.raw == B.g1_compress(want)


def test_affine_batch_add_every_branch(hc):
    """The pair logic of the batched affine additions (csrc/affine_batch.cuh, used by msm_affine.cu): generic sums,
    P + P (doubling slope through the shared inversion), P + (-P), infinity on either or both sides, all inside ONE
    batch with one inversion, against the integer oracle."""
    rnd = random.Random(11)
    inf = B.G1_INF
    pts = [rand_g1(rnd) for _ in range(40)]
    a, b = [], []
    for i in range(16):
        a.append(pts[i]), b.append(pts[16 + i])
    a += [pts[0], pts[1], inf, pts[2], inf, pts[3], B.g1_neg(pts[4])]
    b += [pts[0], B.g1_neg(pts[1]), pts[5], inf, inf, pts[3], pts[4]]
    for _ in range(20):
        p = pts[rnd.randrange(40)]
        q = [pts[rnd.randrange(40)], p, B.g1_neg(p), inf][rnd.randrange(4)]
        a.append(p), b.append(q)
    n = len(a)
    assert n <= 64
    out = C.create_string_buffer(48 * n)
    assert hc.hc_affine_batch_add(out, b"".join(B.g1_compress(x) for x in a), b"".join(B.g1_compress(x) for x in b), n) == 1
    for i in range(n):
        assert out.raw[48 * i : 48 * i + 48] == B.g1_compress(B.g1_add(a[i], b[i])), i
    # a batch of one, and a batch whose pairs are all exceptional
    one = C.create_string_buffer(48)
    assert hc.hc_affine_batch_add(one, B.g1_compress(pts[7]), B.g1_compress(pts[7]), 1) == 1
    assert one.raw == B.g1_compress(B.g1_dbl(pts[7]))
    exc = C.create_string_buffer(48 * 3)
    assert hc.hc_affine_batch_add(exc, b"".join(B.g1_compress(x) for x in (inf, pts[8], pts[9])), b"".join(B.g1_compress(x) for x in (inf, B.g1_neg(pts[8]), inf)), 3) == 1
    assert exc.raw == B.g1_compress(inf) * 2 + B.g1_compress(pts[9])


def test_g1_validate_edge_cases(hc):
    """The encoding edge cases of src/test/tests.c:536-745 (validate_kzg_g1)."""
    rnd = random.Random(6)
    out = C.create_string_buffer(48)
    good = B.g1_compress(rand_g1(rnd))

    def check(b, want_ok):
        got = hc.hc_g1_validate(out, bytes(b))
        assert bool(got) == want_ok, bytes(b).hex()
        if want_ok:
            assert out.raw == bytes(b)

    check(good, True)
    check(B.g1_compress(B.G1_INF), True)
    check(B.g1_compress(B.G1_GEN_J), True)
    b = bytearray(good); b[0] &= 0x7F; check(b, False)  # compressed flag cleared
    b = bytearray(B.g1_compress(B.G1_INF)); b[0] |= 0x20; check(b, False)  # inf with sign bit
    b = bytearray(B.g1_compress(B.G1_INF)); b[47] = 1; check(b, False)  # inf with payload
    b = bytearray(B.g1_compress(B.G1_INF)); b[0] = 0x40; check(b, False)  # inf without compressed bit
    # x >= p
    b = bytearray(P.to_bytes(48, "big")); b[0] |= 0x80; check(b, False)
    b = bytearray((P + 1).to_bytes(48, "big")); b[0] |= 0x80; check(b, False)
    # x = 0
    b = bytearray(48); b[0] = 0x80; check(b, False)
    # not on curve / on curve but not in G1
    n_off = n_out = 0
    x = 5
    while n_off < 3 or n_out < 3:
        x += 1
        y = B.fp_sqrt((x**3 + 4) % P)
        b = bytearray(x.to_bytes(48, "big")); b[0] |= 0x80
        if y is None:
            n_off += 1
            check(b, False)
        else:
            in_g1 = B.g1_in_subgroup((x, y, 1))
            n_out += not in_g1
            check(b, in_g1)
            # uncompress alone accepts it (no subgroup check in load_trusted_setup, setup.c:447)
            assert hc.hc_g1_uncompress(out, bytes(b)) == 1
    # sign bit selects the other root
    p = B.g1_to_affine(rand_g1(rnd))
    for y in (p[1], P - p[1]):
        enc = B.g1_compress((p[0], y, 1))
        assert hc.hc_g1_uncompress(out, enc) == 1 and out.raw == enc


def test_g1_validate_levels_and_basez(hc):
    """The table-producing validation (g1.cuh g1a_validate_levels): same verdict as g1a_validate, levels
    = 2^(8j) P and 2^(8j) [|z|]P; base-|z| digits recompose the scalar; [k]P from the four bases."""
    rnd = random.Random(11)
    out = C.create_string_buffer(48 * 18)
    Z = B.BLS_X
    for trial in range(3):
        p = rand_g1(rnd) if trial else B.G1_GEN_J
        assert hc.hc_g1_validate_levels(out, B.g1_compress(p)) == 1
        q = B.g1_mul(p, Z)
        for j in range(9):
            assert out.raw[48 * j : 48 * j + 48] == B.g1_compress(B.g1_mul(p, 1 << (8 * j))), j
            assert out.raw[48 * (9 + j) : 48 * (10 + j)] == B.g1_compress(B.g1_mul(q, 1 << (8 * j))), j
    # infinity: accepted, all levels infinity
    inf = B.g1_compress(B.G1_INF)
    assert hc.hc_g1_validate_levels(out, inf) == 1 and out.raw == inf * 18
    # on the curve but outside G1: rejected
    x = 5
    while True:
        x += 1
        y = B.fp_sqrt((x**3 + 4) % P)
        if y is not None and not B.g1_in_subgroup((x, y, 1)):
            b = bytearray(x.to_bytes(48, "big")); b[0] |= 0x80
            assert hc.hc_g1_validate_levels(out, bytes(b)) == 0
            break
    # balanced base-|z| expansion (mod r)
    a = (C.c_int64 * 4)()
    for k in EDGE_R + [Z**3 * (Z // 2 + 5) % R, Z // 2, Z // 2 + 1, Z**2 * (Z // 2 + 1)] + [rnd.randrange(R) for _ in range(200)]:
        hc.hc_basez_split(a, limbs(k, 8))
        assert all(abs(int(v)) <= Z // 2 + 1 for v in a), k
        assert sum(int(v) * Z**i for i, v in enumerate(a)) % R == k % R, k
    # [k]P = a0 P + a1 Q - a2 phi(P) - a3 phi(Q) with phi(x, y) = (beta x, y), beta = FP_BETA_A
    beta = pow(2, (P - 1) // 3, P)
    p = rand_g1(rnd)
    pa = B.g1_to_affine(p)
    q = B.g1_mul(p, Z)
    qa = B.g1_to_affine(q)
    k = rnd.randrange(R)
    hc.hc_basez_split(a, limbs(k, 8))
    acc = B.g1_add(B.g1_mul(p, int(a[0]) % R), B.g1_mul(q, int(a[1]) % R))
    acc = B.g1_add(acc, B.g1_neg(B.g1_mul((pa[0] * beta % P, pa[1], 1), int(a[2]) % R)))
    acc = B.g1_add(acc, B.g1_neg(B.g1_mul((qa[0] * beta % P, qa[1], 1), int(a[3]) % R)))
    assert B.g1_eq(acc, B.g1_mul(p, k))


def g2_compress(q):
    """ZCash G2 compression of a Jacobian oracle point (test helper)."""
    a = B.g2_to_affine(q)
    if a is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), (y0, y1) = a
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80
    big = (y1 > (P - 1) // 2) if y1 != 0 else (y0 > (P - 1) // 2)
    if big:
        b[0] |= 0x20
    return bytes(b)


def test_g2_compress_helper_roundtrip():
    q = B.g2_mul(B.G2_GEN_J, 987654321)
    assert B.g2_to_affine(B.g2_uncompress(g2_compress(q))) == B.g2_to_affine(q)


def test_pairing_host(hc):
    rnd = random.Random(7)
    a, b = rnd.randrange(1, R), rnd.randrange(1, R)
    G1, G2 = B.G1_GEN_J, B.G2_GEN_J
    aG1, bG2 = B.g1_mul(G1, a), B.g2_mul(G2, b)
    abG1 = B.g1_mul(G1, a * b % R)
    c = lambda p: B.g1_compress(p)
    # e(aG1, bG2) * e(-abG1, G2) == 1
    assert hc.hc_pairing_product_is_one(c(aG1), g2_compress(bG2), c(B.g1_neg(abG1)), g2_compress(G2)) == 1
    assert hc.hc_pairing_product_is_one(c(aG1), g2_compress(bG2), c(abG1), g2_compress(G2)) == 0
    assert hc.hc_pairing_product_is_one(c(aG1), g2_compress(bG2), c(B.g1_neg(B.g1_mul(G1, (a * b + 1) % R))), g2_compress(G2)) == 0
    # infinity handling: e(inf, Q) = 1
    inf = c(B.G1_INF)
    assert hc.hc_pairing_product_is_one(inf, g2_compress(bG2), inf, g2_compress(G2)) == 1
    assert hc.hc_pairing_product_is_one(c(aG1), g2_compress(bG2), inf, g2_compress(G2)) == 0
    assert hc.hc_cyclotomic_consistency(c(aG1), g2_compress(bG2)) == 1


def test_cooperative_pairing_schedule_host(hc):
    """pairing_coop.cuh (lane schedule from tools/gen_pairing_tables.py) == the serial pairing."""
    rnd = random.Random(8)
    G1, G2 = B.G1_GEN_J, B.G2_GEN_J
    c = lambda p: B.g1_compress(p)
    inf = c(B.G1_INF)
    for scale in (0, 1):
        a, b = rnd.randrange(1, R), rnd.randrange(1, R)
        aG1, bG2, abG1 = B.g1_mul(G1, a), B.g2_mul(G2, b), B.g1_mul(G1, a * b % R)
        # negate_first: e(-abG1, G2) * e(aG1, bG2) == 1
        assert hc.hc_coop_pairing_product_is_one(c(abG1), g2_compress(G2), c(aG1), g2_compress(bG2), 1, scale) == 1
        assert hc.hc_coop_pairing_product_is_one(c(abG1), g2_compress(G2), c(aG1), g2_compress(bG2), 0, scale) == 0
        assert hc.hc_coop_pairing_product_is_one(c(B.g1_neg(abG1)), g2_compress(G2), c(aG1), g2_compress(bG2), 0, scale) == 1
        wrong = B.g1_mul(G1, (a * b + 1) % R)
        assert hc.hc_coop_pairing_product_is_one(c(wrong), g2_compress(G2), c(aG1), g2_compress(bG2), 1, scale) == 0
        assert hc.hc_coop_pairing_product_is_one(inf, g2_compress(G2), inf, g2_compress(bG2), 1, scale) == 1
        assert hc.hc_coop_pairing_product_is_one(inf, g2_compress(G2), c(aG1), g2_compress(bG2), 1, scale) == 0


def test_cooperative_pairing_with_the_radix_29_wide_product():
    """The alternative wide product of pairing_coop.cuh (29-bit limbs, 64-bit column accumulators, R_w = 2^406; off in
    the product build because it measured slower) through the same cooperative pairing check."""
    hc29 = C.CDLL(build(defines=("COOP_W29=1",), tag="_w29"))
    rnd = random.Random(29)
    G1, G2 = B.G1_GEN_J, B.G2_GEN_J
    c = lambda p: B.g1_compress(p)
    for scale in (0, 1):
        a, b = rnd.randrange(1, R), rnd.randrange(1, R)
        aG1, bG2, abG1 = B.g1_mul(G1, a), B.g2_mul(G2, b), B.g1_mul(G1, a * b % R)
        assert hc29.hc_coop_pairing_product_is_one(c(abG1), g2_compress(G2), c(aG1), g2_compress(bG2), 1, scale) == 1
        assert hc29.hc_coop_pairing_product_is_one(c(abG1), g2_compress(G2), c(aG1), g2_compress(bG2), 0, scale) == 0
    inf = c(B.G1_INF)
    assert hc29.hc_coop_pairing_product_is_one(inf, g2_compress(G2), inf, g2_compress(bG2), 1, 0) == 1


def _sim_branches(a, mod):
    """Which rare branches of inv_binary (field.cuh: merged shift steps) the input reaches: a low word of zeros in
    the 'v even' / 'u even' steps, and in the difference of the two subtraction steps."""
    u, v, hit = mod, a, set()
    while v:
        if v % 2 == 0:
            if v % (1 << 32) == 0:
                hit.add("v_even_zero_word")
            t = min(31, (v & -v).bit_length() - 1)
            v >>= t
        elif u % 2 == 0:
            if u % (1 << 32) == 0:
                hit.add("u_even_zero_word")
            t = min(31, (u & -u).bit_length() - 1)
            u >>= t
        elif v >= u:
            d = v - u
            if d and d % (1 << 32) == 0:
                hit.add("v_minus_u_zero_word")
            v = d >> (1 if d == 0 else min(31, (d & -d).bit_length() - 1))
        else:
            d = u - v
            if d % (1 << 32) == 0:
                hit.add("u_minus_v_zero_word")
            u = d >> min(31, (d & -d).bit_length() - 1)
    return hit


def test_inversion_differences_with_a_zero_low_word(hc):
    """Inputs built so that a difference of the binary-Euclid loop has 32 or more trailing zero bits (probability 2^-32
    for random inputs): the merged shift then goes 31 bits at a time and continues in the even steps."""
    for mod, n, fn in ((P, 12, hc.hc_fp_inv), (R, 8, hc.hc_fr_inv)):
        out = (C.c_uint32 * n)()
        cases, seen = [], set()
        for e in (32, 33, 40, 63, 64, 65, 96, 200):
            for m in (1, 3, 5, 0x1234567):
                a = mod - (m << e)  # first step: u - v = m 2^e
                if 0 < a < mod:
                    cases.append(a)
        # second step v - u' with u' = (mod - a) / 2: 3 a = mod (mod 2^k)
        for k in (33, 40, 64, 70):
            a0 = mod * pow(3, -1, 1 << k) % (1 << k)
            for j in range((mod // 2) >> k, ((mod // 2) >> k) + 200):
                a = a0 + (j << k)
                if mod // 3 < a < mod and (mod - a) % 4 == 2:
                    cases.append(a)
                    break
        for a in cases:
            seen |= _sim_branches(a, mod)
            fn(out, limbs(a, n))
            assert val(out) == pow(a, mod - 2, mod), hex(a)
        assert {"u_minus_v_zero_word", "v_minus_u_zero_word", "u_even_zero_word"} <= seen, seen
