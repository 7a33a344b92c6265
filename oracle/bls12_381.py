"""BLS12-381 arithmetic on Python integers -- ORACLE / TEST INFRASTRUCTURE ONLY.

This file is part of `oracle/`: a CPU restatement of what the reference computes, used only by
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg as the *checker*.  Nothing in
the product path (`c-kzg-4844_b200/`) may import it.

It restates the arithmetic the reference delegates to blst (the vendored submodule
/root/reference/blst @ e7f90de), at the level of mathematical definitions rather than limb code:

  * fields / constants ........ blst/src/consts.c:10-36, blst/src/fields.h:14-50
  * G1 formulas ............... blst/src/ec_ops.h:40-340 (Jacobian add/double), blst/src/e1.c
  * G1 (de)serialisation ...... blst/src/e1.c:201-294 (ZCash flags, sgn0), subgroup check
                                blst/src/map_to_g1.c:512-548
  * G2 ........................ blst/src/e2.c
  * tower + pairing ........... blst/src/fp12_tower.c, blst/src/pairing.c:220-261,371-404

Every observable output of the c-kzg API is a canonical encoding (compressed point / canonical
scalar / bool), so parity does not depend on matching blst's internal representations
(SURVEY.md §0.2).  Parity of this file is PINNED by tests/test_oracle_golden.py against the
consensus-spec vectors (tests/golden) and against the compiled reference (oracle/_ref).
"""

# ---------------------------------------------------------------------------------------------
# Constants (blst/src/consts.c:10-36)
# ---------------------------------------------------------------------------------------------
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_X = 0xD201000000010000  # |z|; the curve parameter is z = -BLS_X (consts.c:9)
BLS_X_IS_NEG = True

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (
        0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
    ),
    (
        0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
    ),
)

# ---------------------------------------------------------------------------------------------
# Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - (1+u)), Fp12 = Fp6[w]/(w^2 - v)   (fp12_tower.c)
# ---------------------------------------------------------------------------------------------


def fp_inv(a):
    return pow(a, P - 2, P)


def fp_sqrt(a):
    """p = 3 mod 4 -> candidate a^((p+1)/4) (blst/src/sqrt.c:10-59).  Returns None if non-residue."""
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


F2_ZERO = (0, 0)
F2_ONE = (1, 0)


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return (-a[0] % P, -a[1] % P)


def f2_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return ((a0 * b0 - a1 * b1) % P, (a0 * b1 + a1 * b0) % P)


def f2_sqr(a):
    a0, a1 = a
    return ((a0 + a1) * (a0 - a1) % P, 2 * a0 * a1 % P)


def f2_mul_fp(a, k):
    return (a[0] * k % P, a[1] * k % P)


def f2_conj(a):
    return (a[0], -a[1] % P)


def f2_inv(a):
    a0, a1 = a
    t = fp_inv((a0 * a0 + a1 * a1) % P)
    return (a0 * t % P, -a1 * t % P)


def f2_mul_xi(a):
    """multiply by xi = 1 + u"""
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)


def f2_is_zero(a):
    return a[0] % P == 0 and a[1] % P == 0


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


def f2_sqrt(a):
    """Square root in Fp2 (p = 3 mod 4, Adj-Rodriguez alg. 9); None if `a` is a non-residue."""
    if f2_is_zero(a):
        return F2_ZERO
    a1 = f2_pow(a, (P - 3) // 4)
    alpha = f2_mul(f2_sqr(a1), a)
    x0 = f2_mul(a1, a)
    if alpha == (P - 1, 0):
        cand = f2_mul((0, 1), x0)
    else:
        b = f2_pow(f2_add(F2_ONE, alpha), (P - 1) // 2)
        cand = f2_mul(b, x0)
    return cand if f2_sqr(cand) == (a[0] % P, a[1] % P) else None


F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b):
    return (f2_add(a[0], b[0]), f2_add(a[1], b[1]), f2_add(a[2], b[2]))


def f6_sub(a, b):
    return (f2_sub(a[0], b[0]), f2_sub(a[1], b[1]), f2_sub(a[2], b[2]))


def f6_neg(a):
    return (f2_neg(a[0]), f2_neg(a[1]), f2_neg(a[2]))


def f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    t0, t1, t2 = f2_mul(a0, b0), f2_mul(a1, b1), f2_mul(a2, b2)
    c0 = f2_add(t0, f2_mul_xi(f2_sub(f2_mul(f2_add(a1, a2), f2_add(b1, b2)), f2_add(t1, t2))))
    c1 = f2_add(f2_sub(f2_mul(f2_add(a0, a1), f2_add(b0, b1)), f2_add(t0, t1)), f2_mul_xi(t2))
    c2 = f2_add(f2_sub(f2_mul(f2_add(a0, a2), f2_add(b0, b2)), f2_add(t0, t2)), t1)
    return (c0, c1, c2)


def f6_mul_v(a):
    """multiply by v: (a0,a1,a2) -> (xi*a2, a0, a1)"""
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_inv(a):
    a0, a1, a2 = a
    c0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    c1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    c2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    t = f2_add(f2_mul(a0, c0), f2_mul_xi(f2_add(f2_mul(a2, c1), f2_mul(a1, c2))))
    t = f2_inv(t)
    return (f2_mul(c0, t), f2_mul(c1, t), f2_mul(c2, t))


F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    t0, t1 = f6_mul(a0, b0), f6_mul(a1, b1)
    c1 = f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), f6_add(t0, t1))
    return (f6_add(t0, f6_mul_v(t1)), c1)


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a):
    return (a[0], f6_neg(a[1]))


def f12_inv(a):
    a0, a1 = a
    t = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1))))
    return (f6_mul(a0, t), f6_neg(f6_mul(a1, t)))


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a)
        e >>= 1
    return r


def f12_is_one(a):
    return a == F12_ONE


# Frobenius: coefficients gamma_k = xi^(k (p-1)/6), k = 1..5
_FROB_GAMMA = [f2_pow((1, 1), k * (P - 1) // 6) for k in range(6)]


def f12_frobenius(a):
    """a^p.  Writing a = sum_{k<6} c_k w^k with c_k in Fp2 (w^2 = v): c_k -> conj(c_k)*gamma_k."""
    (c0, c2, c4), (c1, c3, c5) = a
    c = [c0, c1, c2, c3, c4, c5]
    d = [f2_mul(f2_conj(c[k]), _FROB_GAMMA[k]) for k in range(6)]
    return ((d[0], d[2], d[4]), (d[1], d[3], d[5]))


# ---------------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 4, Jacobian (X, Y, Z), infinity <=> Z == 0            (ec_ops.h, e1.c)
# ---------------------------------------------------------------------------------------------
G1_INF = (0, 1, 0)
G1_GEN_J = (G1_GEN[0], G1_GEN[1], 1)


def g1_is_inf(p):
    return p[2] % P == 0


def g1_dbl(p):
    X, Y, Z = p
    if Z == 0 or Y == 0:
        return G1_INF
    A = X * X % P
    B = Y * Y % P
    C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    X3 = (E * E - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def g1_add(p, q):
    """Complete add-or-double, as blst_p1_add_or_double (src/common/ec.c:29)."""
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if Z1 == 0:
        return q
    if Z2 == 0:
        return p
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        return g1_dbl(p) if S1 == S2 else G1_INF
    H = (U2 - U1) % P
    I = 4 * H * H % P
    J = H * I % P
    r = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * S1 * J) % P
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P
    return (X3, Y3, Z3)


def g1_neg(p):
    return (p[0], -p[1] % P, p[2])


def g1_sub(p, q):
    return g1_add(p, g1_neg(q))


def g1_mul(p, k):
    """Scalar multiplication by a non-negative integer (src/common/ec.c:53 -> blst_p1_mult)."""
    acc = G1_INF
    for bit in bin(k)[2:] if k else "":
        acc = g1_dbl(acc)
        if bit == "1":
            acc = g1_add(acc, p)
    return acc


def g1_to_affine(p):
    """Returns None for infinity."""
    X, Y, Z = p
    if Z % P == 0:
        return None
    zi = fp_inv(Z)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g1_eq(p, q):
    return g1_to_affine(p) == g1_to_affine(q)


def g1_on_curve_affine(x, y):
    return (y * y - x * x * x - 4) % P == 0


def g1_in_subgroup(p):
    """Prime-order subgroup membership.  blst uses an endomorphism-based test
    (map_to_g1.c:512-548); the predicate it decides is exactly [r]P == infinity."""
    return g1_is_inf(g1_mul(p, R))


def g1_compress(p):
    """ZCash compressed encoding (blst/src/e1.c:201-225).  bit7 compressed, bit6 infinity,
    bit5 = y is the lexicographically larger root."""
    a = g1_to_affine(p)
    if a is None:
        return bytes([0xC0]) + bytes(47)
    x, y = a
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def g1_uncompress(b):
    """blst_p1_uncompress (blst/src/e1.c:236-294).  Returns a Jacobian point or None on any
    encoding error.  No subgroup check here (validate_kzg_g1 adds it, src/common/bytes.c:81)."""
    if len(b) != 48:
        return None
    if not b[0] & 0x80:
        return None  # uncompressed form is not accepted by the 48-byte API
    if b[0] & 0x40:  # infinity: every other bit must be clear
        if b[0] & 0x3F or any(b[1:]):
            return None
        return G1_INF
    sign = bool(b[0] & 0x20)
    x = int.from_bytes(bytes([b[0] & 0x1F]) + bytes(b[1:]), "big")
    if x >= P or x == 0:  # x == 0: (0,+-2) -> BLST_POINT_NOT_IN_GROUP (e1.c:289)
        return None
    y = fp_sqrt((x * x * x + 4) % P)
    if y is None:
        return None
    if (y > (P - 1) // 2) != sign:
        y = P - y
    return (x, y, 1)


# ---------------------------------------------------------------------------------------------
# G2 on the twist y^2 = x^3 + 4(1+u), Jacobian over Fp2                       (e2.c)
# ---------------------------------------------------------------------------------------------
G2_INF = (F2_ZERO, F2_ONE, F2_ZERO)
G2_GEN_J = (G2_GEN[0], G2_GEN[1], F2_ONE)
B2 = (4, 4)


def g2_is_inf(p):
    return f2_is_zero(p[2])


def g2_dbl(p):
    X, Y, Z = p
    if f2_is_zero(Z) or f2_is_zero(Y):
        return G2_INF
    A = f2_sqr(X)
    B = f2_sqr(Y)
    C = f2_sqr(B)
    t = f2_sub(f2_sub(f2_sqr(f2_add(X, B)), A), C)
    D = f2_add(t, t)
    E = f2_add(f2_add(A, A), A)
    X3 = f2_sub(f2_sqr(E), f2_add(D, D))
    C8 = f2_mul_fp(C, 8)
    Y3 = f2_sub(f2_mul(E, f2_sub(D, X3)), C8)
    YZ = f2_mul(Y, Z)
    return (X3, Y3, f2_add(YZ, YZ))


def g2_add(p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if f2_is_zero(Z1):
        return q
    if f2_is_zero(Z2):
        return p
    Z1Z1 = f2_sqr(Z1)
    Z2Z2 = f2_sqr(Z2)
    U1 = f2_mul(X1, Z2Z2)
    U2 = f2_mul(X2, Z1Z1)
    S1 = f2_mul(f2_mul(Y1, Z2), Z2Z2)
    S2 = f2_mul(f2_mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        return g2_dbl(p) if S1 == S2 else G2_INF
    H = f2_sub(U2, U1)
    I = f2_mul_fp(f2_sqr(H), 4)
    J = f2_mul(H, I)
    r = f2_mul_fp(f2_sub(S2, S1), 2)
    V = f2_mul(U1, I)
    X3 = f2_sub(f2_sub(f2_sqr(r), J), f2_add(V, V))
    SJ = f2_mul(S1, J)
    Y3 = f2_sub(f2_mul(r, f2_sub(V, X3)), f2_add(SJ, SJ))
    Z3 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(Z1, Z2)), Z1Z1), Z2Z2), H)
    return (X3, Y3, Z3)


def g2_neg(p):
    return (p[0], f2_neg(p[1]), p[2])


def g2_mul(p, k):
    acc = G2_INF
    for bit in bin(k)[2:] if k else "":
        acc = g2_dbl(acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc


def g2_to_affine(p):
    X, Y, Z = p
    if f2_is_zero(Z):
        return None
    zi = f2_inv(Z)
    zi2 = f2_sqr(zi)
    return (f2_mul(X, zi2), f2_mul(Y, f2_mul(zi2, zi)))


def g2_uncompress(b):
    """blst_p2_uncompress (blst/src/e2.c): 96 bytes = x.c1 || x.c0 big-endian, flags in byte 0.
    The sign flag is that of y.c1, or of y.c0 when y.c1 == 0."""
    if len(b) != 96 or not b[0] & 0x80:
        return None
    if b[0] & 0x40:
        if b[0] & 0x3F or any(b[1:]):
            return None
        return G2_INF
    sign = bool(b[0] & 0x20)
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + bytes(b[1:48]), "big")
    x0 = int.from_bytes(b[48:96], "big")
    if x0 >= P or x1 >= P:
        return None
    x = (x0, x1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B2))
    if y is None:
        return None
    big = (y[1] > (P - 1) // 2) if y[1] != 0 else (y[0] > (P - 1) // 2)
    if big != sign:
        y = f2_neg(y)
    return (x, y, F2_ONE)


# ---------------------------------------------------------------------------------------------
# Optimal-ate pairing check                                   (pairing.c:220-261, 371-404)
# ---------------------------------------------------------------------------------------------


def _line_to_f12(c00, c01, c11):
    """sparse element  c00 + c01*v + c11*v*w  (see derivation in DESIGN.md: the M-twist untwist
    (x',y') -> (x'/w^2, y'/w^3) puts a line, scaled by w^3, at exactly these three slots)."""
    return ((c00, c01, F2_ZERO), (F2_ZERO, c11, F2_ZERO))


def miller_loop(Pa, Qa):
    """f_{|z|,Q}(P), conjugated for z<0.  Pa affine G1 (x,y) or None, Qa affine G2 or None."""
    if Pa is None or Qa is None:
        return F12_ONE
    xP, yP = Pa
    xQ, yQ = Qa
    xT, yT = xQ, yQ
    f = F12_ONE
    for bit in bin(BLS_X)[3:]:
        # tangent at T
        lam = f2_mul(f2_mul_fp(f2_sqr(xT), 3), f2_inv(f2_add(yT, yT)))
        line = _line_to_f12(f2_sub(f2_mul(lam, xT), yT), f2_neg(f2_mul_fp(lam, xP)), (yP, 0))
        f = f12_mul(f12_sqr(f), line)
        x3 = f2_sub(f2_sqr(lam), f2_add(xT, xT))
        yT = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT)
        xT = x3
        if bit == "1":
            lam = f2_mul(f2_sub(yT, yQ), f2_inv(f2_sub(xT, xQ)))
            line = _line_to_f12(f2_sub(f2_mul(lam, xT), yT), f2_neg(f2_mul_fp(lam, xP)), (yP, 0))
            f = f12_mul(f, line)
            x3 = f2_sub(f2_sub(f2_sqr(lam), xT), xQ)
            yT = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT)
            xT = x3
    return f12_conj(f) if BLS_X_IS_NEG else f


_HARD_EXP = (P**4 - P**2 + 1) // R


def final_exp(f):
    """f^((p^12-1)/r): easy part (p^6-1)(p^2+1), then the hard part by plain exponentiation."""
    f = f12_mul(f12_conj(f), f12_inv(f))
    f = f12_mul(f12_frobenius(f12_frobenius(f)), f)
    return f12_pow(f, _HARD_EXP)


def pairings_verify(a1, a2, b1, b2):
    """e(a1,a2) == e(b1,b2), evaluated as e(-a1,a2)*e(b1,b2) == 1  (src/common/utils.c:172-196).
    a1,b1 Jacobian G1; a2,b2 Jacobian G2."""
    f = f12_mul(
        miller_loop(g1_to_affine(g1_neg(a1)), g2_to_affine(a2)),
        miller_loop(g1_to_affine(b1), g2_to_affine(b2)),
    )
    return f12_is_one(final_exp(f))
