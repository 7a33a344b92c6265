#!/usr/bin/env python3
"""Generate c-kzg-4844_b200/csrc/pairing_tables.cuh: lane schedules for the thread-cooperative Fp12
arithmetic of the pairing check (csrc/pairing_coop.cuh).

Every tower operation the Miller loop and the final exponentiation need (Fp12 product, Fp12 square,
product with a sparse line, cyclotomic square) is BILINEAR of depth one:

    out_k = sum_L  c_kL * ( X_L(inputs) * Y_L(inputs) )         k = 0..11 (Fp coefficients)

with X_L, Y_L, and the output combinations small signed sums.  This script runs the usual Karatsuba /
complex-squaring / Granger-Scott formulas on symbolic linear forms, records the products, and emits the
three term lists per operation.  On the GPU lane L evaluates X_L, Y_L and multiplies them (one
Montgomery product per lane, 18..54 lanes busy), then 12 lanes gather the outputs.

Coefficient order of an Fp12 element (12 Fp): index 2k+j = j-th component (0 real, 1 imaginary) of the
Fp2 coefficient of w^k, k = 0..5  (w^2 = v, so w^0,w^2,w^4 = c0.{c0,c1,c2} and w^1,w^3,w^5 = c1.{c0,c1,c2}).

The tables are checked here against the integer tower arithmetic of oracle/bls12_381.py before being
written.  Run: python tools/gen_pairing_tables.py
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bls12_381 as B  # noqa: E402

P = B.P


# ---- symbolic linear forms -----------------------------------------------------------------------
class Form:
    """integer linear combination of atoms (input coefficients or products)"""

    __slots__ = ("t",)

    def __init__(self, t=None):
        self.t = dict(t or {})

    def __add__(self, o):
        r = dict(self.t)
        for k, v in o.t.items():
            r[k] = r.get(k, 0) + v
            if r[k] == 0:
                del r[k]
        return Form(r)

    def __neg__(self):
        return Form({k: -v for k, v in self.t.items()})

    def __sub__(self, o):
        return self + (-o)

    def dbl(self):
        return self + self


class Recorder:
    def __init__(self):
        self.products = []  # (Xform, Yform)

    def mul(self, x, y):
        self.products.append((x, y))
        return Form({("p", len(self.products) - 1): 1})


REC = None


def fmul(x, y):
    return REC.mul(x, y)


# Fp2 on forms
def f2_add(a, b): return (a[0] + b[0], a[1] + b[1])
def f2_sub(a, b): return (a[0] - b[0], a[1] - b[1])
def f2_dbl(a): return (a[0].dbl(), a[1].dbl())
def f2_mul_xi(a): return (a[0] - a[1], a[0] + a[1])


def f2_mul(a, b):
    t0, t1 = fmul(a[0], b[0]), fmul(a[1], b[1])
    t2 = fmul(a[0] + a[1], b[0] + b[1])
    return (t0 - t1, t2 - t0 - t1)


def f2_sqr(a):
    s = fmul(a[0] + a[1], a[0] - a[1])
    m = fmul(a[0], a[1])
    return (s, m.dbl())


def f2_mul_fp_real(a, b_real):
    """a * (b_real + 0 u)"""
    return (fmul(a[0], b_real), fmul(a[1], b_real))


def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_mul_v(a): return (f2_mul_xi(a[2]), a[0], a[1])


def f6_mul(a, b):
    t0, t1, t2 = f2_mul(a[0], b[0]), f2_mul(a[1], b[1]), f2_mul(a[2], b[2])
    c0 = f2_add(t0, f2_mul_xi(f2_sub(f2_mul(f2_add(a[1], a[2]), f2_add(b[1], b[2])), f2_add(t1, t2))))
    c1 = f2_add(f2_sub(f2_mul(f2_add(a[0], a[1]), f2_add(b[0], b[1])), f2_add(t0, t1)), f2_mul_xi(t2))
    c2 = f2_add(f2_sub(f2_mul(f2_add(a[0], a[2]), f2_add(b[0], b[2])), f2_add(t0, t2)), t1)
    return (c0, c1, c2)


def f6_mul_by_01(a, b0, b1):
    t0, t1 = f2_mul(a[0], b0), f2_mul(a[1], b1)
    c0 = f2_add(t0, f2_mul_xi(f2_sub(f2_mul(f2_add(a[1], a[2]), b1), t1)))
    c1 = f2_sub(f2_sub(f2_mul(f2_add(a[0], a[1]), f2_add(b0, b1)), t0), t1)
    c2 = f2_add(f2_sub(f2_mul(f2_add(a[0], a[2]), b0), t0), t1)
    return (c0, c1, c2)


def f6_mul_by_1_real(a, c_real):
    """a * (C v) with C = (c_real, 0)"""
    return (f2_mul_xi(f2_mul_fp_real(a[2], c_real)), f2_mul_fp_real(a[0], c_real), f2_mul_fp_real(a[1], c_real))


def f12_mul(a, b):
    t0, t1 = f6_mul(a[0], b[0]), f6_mul(a[1], b[1])
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a[0], a[1]), f6_add(b[0], b[1])), t0), t1)
    return (f6_add(t0, f6_mul_v(t1)), c1)


def f12_sqr(a):
    t = f6_mul(a[0], a[1])
    s = f6_mul(f6_add(a[0], a[1]), f6_add(a[0], f6_mul_v(a[1])))
    return (f6_sub(f6_sub(s, t), f6_mul_v(t)), f6_add(t, t))


def f12_mul_by_line(f, A, Bc, c_real):
    """f * (A + Bc v + C v w), C = (c_real, 0).  The (f0+f1)*(A, Bc + C) term needs Bc + C as an Fp2."""
    t0 = f6_mul_by_01(f[0], A, Bc)
    t1 = f6_mul_by_1_real(f[1], c_real)
    BC = (Bc[0] + c_real, Bc[1])
    c1 = f6_sub(f6_sub(f6_mul_by_01(f6_add(f[0], f[1]), A, BC), t0), t1)
    return (f6_add(t0, f6_mul_v(t1)), c1)


def fp4_sqr(a, b):
    t0, t1 = f2_sqr(a), f2_sqr(b)
    o0 = f2_add(f2_mul_xi(t1), t0)
    o1 = f2_sub(f2_sub(f2_sqr(f2_add(a, b)), t0), t1)
    return o0, o1


def f12_cyc_sqr(f):
    z0, z4, z3 = f[0]
    z2, z1, z5 = f[1]
    t0, t1 = fp4_sqr(z0, z1)
    z0 = f2_add(f2_dbl(f2_sub(t0, z0)), t0)
    z1 = f2_add(f2_dbl(f2_add(t1, z1)), t1)
    t0, t1 = fp4_sqr(z2, z3)
    t2, t3 = fp4_sqr(z4, z5)
    z4 = f2_add(f2_dbl(f2_sub(t0, z4)), t0)
    z5 = f2_add(f2_dbl(f2_add(t1, z5)), t1)
    t0 = f2_mul_xi(t3)
    z2 = f2_add(f2_dbl(f2_add(t0, z2)), t0)
    z3 = f2_add(f2_dbl(f2_sub(t2, z3)), t2)
    return ((z0, z4, z3), (z2, z1, z5))


# ---- coefficient order -----------------------------------------------------------------------------
def tower_from_flat(c):
    """12 items (index 2k+j) -> ((c0.c0,c0.c1,c0.c2),(c1.c0,c1.c1,c1.c2)) of Fp2 pairs"""
    k = lambda i: (c[2 * i], c[2 * i + 1])
    return ((k(0), k(2), k(4)), (k(1), k(3), k(5)))


def flat_from_tower(t):
    out = [None] * 12
    for half in (0, 1):
        for i in range(3):
            kk = 2 * i + half
            out[2 * kk], out[2 * kk + 1] = t[half][i]
    return out


def atoms(prefix, n, base):
    return [Form({("in", base + i): 1}) for i in range(n)]


def build(op):
    """-> (products, outputs): products = list of (Xform, Yform) over input atoms; outputs = 12 forms
    over ('p', L) atoms and, for the cyclotomic square, input atoms too."""
    global REC
    REC = Recorder()
    a = atoms("a", 12, 0)
    if op == "mul":
        b = atoms("b", 12, 12)
        out = f12_mul(tower_from_flat(a), tower_from_flat(b))
    elif op == "sqr":
        out = f12_sqr(tower_from_flat(a))
    elif op == "cyc":
        # The Granger-Scott outputs are 3 t +- 2 z with z an INPUT coefficient.  The wide-domain
        # arithmetic never reduces sums, so a value that is copied from input to output would grow
        # geometrically over the 63 consecutive squarings of an exponentiation; route every such
        # term through a product with the constant one (input index 12) instead: products come out
        # of the Montgomery multiplier reduced.
        one = Form({("in", 12): 1})
        out = flat_from_tower(f12_cyc_sqr(tower_from_flat(a)))
        via = {}
        fixed = []
        for o in out:
            f = Form()
            for (k, idx), c in o.t.items():
                if k == "in":
                    if idx not in via:
                        via[idx] = fmul(a[idx], one)
                    f = f + Form({key: v * c for key, v in via[idx].t.items()})
                else:
                    f = f + Form({(k, idx): c})
            fixed.append(f)
        return REC.products, fixed
    elif op == "line":
        # b inputs: 12..13 = A, 14..15 = Bc (already multiplied by the x scaling), 16 = C real part
        b = atoms("b", 5, 12)
        out = f12_mul_by_line(tower_from_flat(a), (b[0], b[1]), (b[2], b[3]), b[4])
    else:
        raise KeyError(op)
    return REC.products, flat_from_tower(out)


def expand(form, kind):
    """signed unit terms: +(idx+1) / -(idx+1); inputs and products share the index space by `kind`"""
    terms = []
    for (k, idx), c in sorted(form.t.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        code = idx + 1 if k == kind else None
        if code is None:
            # cyclotomic square outputs mix products and inputs: inputs are encoded after the products
            assert kind == "p" and k == "in"
            code = 64 + idx + 1
        terms += [code if c > 0 else -code] * abs(c)
    return terms


# ---- numeric check against the oracle --------------------------------------------------------------
def eval_tables(products, outputs, inputs):
    def ev(form, prods):
        s = 0
        for (k, idx), c in form.t.items():
            s += c * (inputs[idx] if k == "in" else prods[idx])
        return s % P
    prods = [ev(x, None) * ev(y, None) % P for x, y in products]
    return [ev(o, prods) for o in outputs]


def oracle_flat(x):
    return flat_from_tower(x)  # oracle elements have the same nested shape


def rand_f12(rnd):
    return tuple(tuple((rnd.randrange(P), rnd.randrange(P)) for _ in range(3)) for _ in range(2))


def flatten_ints(t):
    out = []
    for fp2 in oracle_flat(t):
        pass
    return out


def to_int_list(t):
    fl = flat_from_tower(t)
    return [int(x) for x in fl]


def check():
    rnd = random.Random(5)
    for op in ("mul", "sqr", "line", "cyc"):
        products, outputs = build(op)
        a = rand_f12(rnd)
        ai = to_int_list(a)
        if op == "mul":
            b = rand_f12(rnd)
            got = eval_tables(products, outputs, ai + to_int_list(b))
            want = to_int_list(B.f12_mul(a, b))
        elif op == "sqr":
            got = eval_tables(products, outputs, ai)
            want = to_int_list(B.f12_mul(a, a))
        elif op == "line":
            A, Bc, cr = (rnd.randrange(P), rnd.randrange(P)), (rnd.randrange(P), rnd.randrange(P)), rnd.randrange(P)
            line = ((A, Bc, (0, 0)), ((0, 0), (cr, 0), (0, 0)))
            got = eval_tables(products, outputs, ai + [A[0], A[1], Bc[0], Bc[1], cr])
            want = to_int_list(B.f12_mul(a, line))
        else:
            # cyclotomic subgroup element: easy part of the final exponentiation of a random element
            f = B.f12_mul(B.f12_conj(a), B.f12_inv(a))
            f = B.f12_mul(B.f12_frobenius(B.f12_frobenius(f)), f)
            got = eval_tables(products, outputs, to_int_list(f) + [1])
            want = to_int_list(B.f12_mul(f, f))
        assert got == want, op
        print("%-5s %2d products  ok" % (op, len(products)), file=sys.stderr)


def emit():
    out = []
    out.append("// GENERATED by tools/gen_pairing_tables.py -- do not edit.\n")
    out.append("// Lane schedules for the cooperative Fp12 arithmetic (see the generator's docstring).\n")
    out.append("// Term code t: |t|-1 = operand index, sign = add/subtract.  X/Y forms index the inputs\n")
    out.append("// (0..11 = a, 12.. = b; for the cyclotomic square b[0] is the constant one); output forms index the products.\n")
    out.append("#pragma once\n#include <stdint.h>\n\n")
    out.append("struct CoopOpTable {\n    int nprod;\n    const int16_t* x_off;  // nprod + 1\n    const int16_t* y_off;\n    const int16_t* o_off;  // 13\n    const int8_t* x_terms;\n    const int8_t* y_terms;\n    const int8_t* o_terms;\n};\n\n")
    for op in ("mul", "sqr", "line", "cyc"):
        products, outputs = build(op)
        xs, ys, os_ = [], [], []
        xo, yo, oo = [0], [0], [0]
        for x, y in products:
            xs += expand(x, "in")
            xo.append(len(xs))
            ys += expand(y, "in")
            yo.append(len(ys))
        for o in outputs:
            os_ += expand(o, "p")
            oo.append(len(os_))
        assert max(abs(t) for t in xs + ys) < 64 and max(abs(t) for t in os_) <= 64  # outputs: products only
        # pairing_coop.cuh keeps a negated copy of every `a` coefficient and of every product, never of a `b` operand
        # (line values, the constant one): a subtracted input must be one of the twelve `a` coefficients
        assert all(t > 0 or -t <= 12 for t in xs + ys), op
        # bounds the unreduced sums of pairing_coop.cuh rely on
        assert max(b - a for a, b in zip(xo, xo[1:])) <= 8 and max(b - a for a, b in zip(yo, yo[1:])) <= 8
        assert max(b - a for a, b in zip(oo, oo[1:])) <= 40 and len(products) <= 64
        U = op.upper()
        def arr(ctype, name, vals):
            return "KZG_CONST %s %s[%d] = {%s};\n" % (ctype, name, len(vals), ", ".join(str(v) for v in vals))
        out.append("// ---- %s: %d products ----\n" % (op, len(products)))
        out.append("#define COOP_%s_NPROD %d\n" % (U, len(products)))
        out.append(arr("int16_t", "COOP_%s_XOFF" % U, xo))
        out.append(arr("int16_t", "COOP_%s_YOFF" % U, yo))
        out.append(arr("int16_t", "COOP_%s_OOFF" % U, oo))
        out.append(arr("int8_t", "COOP_%s_XT" % U, xs))
        out.append(arr("int8_t", "COOP_%s_YT" % U, ys))
        out.append(arr("int8_t", "COOP_%s_OT" % U, os_))
        out.append("\n")
        print("%-5s terms: X %d  Y %d  O %d  (max out form %d)" % (op, len(xs), len(ys), len(os_), max(b - a for a, b in zip(oo, oo[1:]))), file=sys.stderr)
    # ---- cyclotomic square, second form (coop_cyc2): every product carries ONE weight (2, 3 or 6: the Granger-Scott
    # outputs are 3 t +- 2 z), applied by the lane that owns the product, so the output forms are plain signed sums
    # of at most 7 scaled products -- one pass of 12 lanes instead of 60 partial sums and a second gathering pass ----
    products, outputs = build("cyc")
    weight = {}
    for o in outputs:
        for (k, idx), c in o.t.items():
            assert k == "p"
            assert weight.setdefault(idx, abs(c)) == abs(c), "a product with two weights"
    W = [weight[i] for i in range(len(products))]
    assert all(w in (2, 3, 6) for w in W)
    o2, oo2, worst = [], [0], 0
    for o in outputs:
        terms = sorted(o.t.items(), key=lambda kv: kv[0][1])
        o2 += [(idx + 1) if c > 0 else -(idx + 1) for (k, idx), c in terms]
        oo2.append(len(o2))
        # bound in multiples of p: a scaled product is < w p, its stored negative 8 p - v <= 8 p
        worst = max(worst, sum(W[idx] if c > 0 else 8 for (k, idx), c in terms))
    assert max(b - a for a, b in zip(oo2, oo2[1:])) <= 8 and worst <= 64, worst
    out.append("// ---- cyc, second form: weights per product, unit output terms (at most 8 per coefficient, sum <= %d p) ----\n" % worst)
    out.append("KZG_CONST int8_t COOP_CYC2_W[%d] = {%s};\n" % (len(W), ", ".join(map(str, W))))
    out.append("KZG_CONST int16_t COOP_CYC2_OOFF[13] = {%s};\n" % ", ".join(map(str, oo2)))
    out.append("KZG_CONST int8_t COOP_CYC2_OT[%d] = {%s};\n\n" % (len(o2), ", ".join(map(str, o2))))
    print("cyc2  weights %s  O %d (max out form %d, bound %d p)" % (sorted(set(W)), len(o2), max(b - a for a, b in zip(oo2, oo2[1:])), worst), file=sys.stderr)
    # ---- constants of the 14-limb "wide" Montgomery domain (R_w = 2^448) ----
    def limbs14(v):
        assert 0 <= v < 1 << 448
        return ", ".join("0x%08xu" % ((v >> (32 * i)) & 0xFFFFFFFF) for i in range(14))
    out.append("// ---- wide domain: 14 limbs, R_w = 2^448; sums are never reduced, only the multiplier reduces ----\n")
    for name, v in (("FPW_MOD", P), ("FPW_ONE", (1 << 448) % P), ("FPW_R2", (1 << 896) % P), ("FPW_C512", (1 << 512) % P),
                    ("FPW_C576", (1 << 576) % P), ("FPW_OFF8", 8 * P), ("FPW_OFF16", 16 * P), ("FPW_OFF64", 64 * P), ("FPW_OFF256", 256 * P), ("FPW_OFF2048", 2048 * P)):
        out.append("KZG_CONST uint32_t %s[14] = {%s};\n" % (name, limbs14(v)))
    # ---- the radix-2^29 multiplier of the wide domain (w_mul29): R_w = 2^406 = 2^(29 * 14) ----
    # conversions between the 12-limb Montgomery form (v 2^384) and the wide form (v 2^RW) go through one wide product
    # with 2^(2 RW - 384); the scaled point coordinates meet a 12-limb line coefficient in one wide product and are
    # carried as s 2^(2 RW - 384), i.e. one product of the 12-limb form with 2^(3 RW - 768)
    RW = 406
    out.append("// ---- radix-2^29 multiplier: R_w = 2^406; products come out < p + 2^376 (not reduced below p) ----\n")
    for name, v in (("FPW29_ONE", (1 << RW) % P), ("FPW29_CIN", (1 << (2 * RW - 384)) % P), ("FPW29_CPT", (1 << (3 * RW - 768)) % P), ("FPW_2P", 2 * P),
                    ("FPW_OFF128", 128 * P)):
        out.append("KZG_CONST uint32_t %s[14] = {%s};\n" % (name, limbs14(v)))
    out.append("KZG_CONST uint32_t FPW_P29[14] = {%s};\n" % ", ".join("0x%08xu" % ((P >> (29 * i)) & ((1 << 29) - 1)) for i in range(14)))
    out.append("#define FPW_INV29 0x%08xu  // -1 / p mod 2^29\n" % ((-pow(P, -1, 1 << 29)) % (1 << 29)))
    # worst-case output sums when a product is < 1.05 p and its stored negative 2 p - v <= 2 p (in p / 100)
    for op in ("mul", "sqr", "line"):
        products, outputs = build(op)
        worst = max(sum(105 * c if c > 0 else 200 * (-c) for c in o.t.values()) for o in outputs)
        print("%-5s worst output sum with unreduced products: %.1f p" % (op, worst / 100), file=sys.stderr)
        assert worst <= 128 * 100
    path = os.path.join(ROOT, "c-kzg-4844_b200", "csrc", "pairing_tables.cuh")
    with open(path, "w") as f:
        f.write("".join(out))
    print("wrote", path, file=sys.stderr)


if __name__ == "__main__":
    check()
    emit()
