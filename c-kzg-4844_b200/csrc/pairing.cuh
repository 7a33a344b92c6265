// Extension tower, G2 and the pairing check for the verify_* paths.
//
// Replaces, for the GPU engine: blst/src/fp12_tower.c (Fp2/Fp6/Fp12 arithmetic), blst/src/e2.c
// (G2 decompression, blst_p2_uncompress), blst/src/pairing.c:220-261 (Miller loop, here in the
// "precomputed lines" form blst also offers at :272-335) and :371-404 (final exponentiation), as used
// by pairings_verify (src/common/utils.c:172-196).
//
// Tower: Fp2 = Fp[u]/(u^2+1); Fp6 = Fp2[v]/(v^3 - xi), xi = 1+u; Fp12 = Fp6[w]/(w^2 - v).
// Every pairing the KZG verifiers need has one of three FIXED G2 arguments (the G2 generator,
// [tau]G2, [tau^64]G2): verify_kzg_proof's variable point [tau]G2 - [z]G2 (src/eip4844/eip4844.c:355)
// is avoided through the equivalent equation e(C - [y]G1 + [z]pi, G2) == e(pi, [tau]G2) (bilinearity;
// pi is subgroup-checked, so the booleans coincide).  Hence all G2 work (65 decompressions, line
// coefficients of the 3 fixed points) happens once at setup time; a verification is
// 2 x (63 + 5) line evaluations + 63 Fp12 squarings + one final exponentiation.
//
// A line through T with slope lambda on the M-twist, evaluated at P = (xP, yP) and scaled by w^3
// (an Fp4 element, killed by the final exponentiation), is the sparse element
//     (lambda*xT - yT)  +  (-lambda * xP) v  +  (yP) v w
// so a precomputed line is the Fp2 pair (A, B) = (lambda*xT - yT, -lambda).
//
// The boolean e(a1,a2) == e(b1,b2) does not depend on which non-degenerate power of the pairing is
// used; the hard part below computes the cube of the canonical reduced pairing
// (3(p^4-p^2+1)/r = (z-1)^2 (z+p)(z^2+p^2-1) + 3), which is 1 exactly when the pairing is.
#pragma once
#include "g1.cuh"

namespace kzg {

// ------------------------------------------------------------------------------------------------
// Fp2
// ------------------------------------------------------------------------------------------------
struct Fp2 {
    Fp c0, c1;
};
KZG_HD Fp2 f2_zero() { Fp2 r; r.c0 = Fp::zero(); r.c1 = Fp::zero(); return r; }
KZG_HD Fp2 f2_one() { Fp2 r; r.c0 = Fp::one(); r.c1 = Fp::zero(); return r; }
KZG_HD bool f2_is_zero(const Fp2& a) { return is_zero(a.c0) && is_zero(a.c1); }
KZG_HD bool f2_eq(const Fp2& a, const Fp2& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
KZG_HD Fp2 f2_add(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); return r; }
KZG_HD Fp2 f2_sub(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); return r; }
KZG_HD Fp2 f2_neg(const Fp2& a) { Fp2 r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
KZG_HD Fp2 f2_dbl(const Fp2& a) { return f2_add(a, a); }
KZG_HD Fp2 f2_conj(const Fp2& a) { Fp2 r; r.c0 = a.c0; r.c1 = neg(a.c1); return r; }
KZG_HD Fp2 f2_mul_xi(const Fp2& a) { Fp2 r; r.c0 = sub(a.c0, a.c1); r.c1 = add(a.c0, a.c1); return r; }
KZG_HD Fp2 f2_mul_fp(const Fp2& a, const Fp& k) { Fp2 r; r.c0 = mul(a.c0, k); r.c1 = mul(a.c1, k); return r; }

KZG_HD_NOINLINE Fp2 f2_mul(const Fp2& a, const Fp2& b) {
    Fp t0 = mul(a.c0, b.c0);
    Fp t1 = mul(a.c1, b.c1);
    Fp t2 = mul(add(a.c0, a.c1), add(b.c0, b.c1));
    Fp2 r;
    r.c0 = sub(t0, t1);
    r.c1 = sub(sub(t2, t0), t1);
    return r;
}
KZG_HD_NOINLINE Fp2 f2_sqr(const Fp2& a) {
    Fp s = add(a.c0, a.c1), d = sub(a.c0, a.c1);
    Fp m = mul(a.c0, a.c1);
    Fp2 r;
    r.c0 = mul(s, d);
    r.c1 = dbl(m);
    return r;
}
KZG_HD_NOINLINE Fp2 f2_inv(const Fp2& a) {
    Fp t = fp_inv(add(sqr(a.c0), sqr(a.c1)));
    Fp2 r;
    r.c0 = mul(a.c0, t);
    r.c1 = neg(mul(a.c1, t));
    return r;
}
template <int NL>
KZG_HD_NOINLINE Fp2 f2_pow_limbs(const Fp2& a, const uint32_t* e) {
    Fp2 r = f2_one();
    for (int i = NL - 1; i >= 0; i--) {
        uint32_t w = e[i];
        for (int b = 31; b >= 0; b--) {
            r = f2_sqr(r);
            if ((w >> b) & 1u) r = f2_mul(r, a);
        }
    }
    return r;
}
// square root in Fp2 for p = 3 mod 4 (Adj & Rodriguez-Henriquez, alg. 9); false if non-residue
KZG_HD_NOINLINE bool f2_sqrt(Fp2& out, const Fp2& a) {
    if (f2_is_zero(a)) {
        out = f2_zero();
        return true;
    }
    Fp2 a1 = f2_pow_limbs<12>(a, FP_P_MINUS_3_DIV_4);
    Fp2 alpha = f2_mul(f2_sqr(a1), a);
    Fp2 x0 = f2_mul(a1, a);
    Fp2 minus_one;
    minus_one.c0 = neg(Fp::one());
    minus_one.c1 = Fp::zero();
    Fp2 cand;
    if (f2_eq(alpha, minus_one)) {
        Fp2 i;
        i.c0 = Fp::zero();
        i.c1 = Fp::one();
        cand = f2_mul(i, x0);
    } else {
        Fp2 b = f2_pow_limbs<12>(f2_add(f2_one(), alpha), FP_P_MINUS_1_DIV_2);
        cand = f2_mul(b, x0);
    }
    out = cand;
    return f2_eq(f2_sqr(cand), a);
}

// ------------------------------------------------------------------------------------------------
// Fp6
// ------------------------------------------------------------------------------------------------
struct Fp6 {
    Fp2 c0, c1, c2;
};
KZG_HD Fp6 f6_zero() { Fp6 r; r.c0 = f2_zero(); r.c1 = f2_zero(); r.c2 = f2_zero(); return r; }
KZG_HD Fp6 f6_one() { Fp6 r; r.c0 = f2_one(); r.c1 = f2_zero(); r.c2 = f2_zero(); return r; }
KZG_HD Fp6 f6_add(const Fp6& a, const Fp6& b) { Fp6 r; r.c0 = f2_add(a.c0, b.c0); r.c1 = f2_add(a.c1, b.c1); r.c2 = f2_add(a.c2, b.c2); return r; }
KZG_HD Fp6 f6_sub(const Fp6& a, const Fp6& b) { Fp6 r; r.c0 = f2_sub(a.c0, b.c0); r.c1 = f2_sub(a.c1, b.c1); r.c2 = f2_sub(a.c2, b.c2); return r; }
KZG_HD Fp6 f6_neg(const Fp6& a) { Fp6 r; r.c0 = f2_neg(a.c0); r.c1 = f2_neg(a.c1); r.c2 = f2_neg(a.c2); return r; }
KZG_HD Fp6 f6_mul_v(const Fp6& a) { Fp6 r; r.c0 = f2_mul_xi(a.c2); r.c1 = a.c0; r.c2 = a.c1; return r; }
KZG_HD bool f6_eq(const Fp6& a, const Fp6& b) { return f2_eq(a.c0, b.c0) && f2_eq(a.c1, b.c1) && f2_eq(a.c2, b.c2); }

KZG_HD_NOINLINE Fp6 f6_mul(const Fp6& a, const Fp6& b) {
    Fp2 t0 = f2_mul(a.c0, b.c0), t1 = f2_mul(a.c1, b.c1), t2 = f2_mul(a.c2, b.c2);
    Fp6 r;
    r.c0 = f2_add(t0, f2_mul_xi(f2_sub(f2_mul(f2_add(a.c1, a.c2), f2_add(b.c1, b.c2)), f2_add(t1, t2))));
    r.c1 = f2_add(f2_sub(f2_mul(f2_add(a.c0, a.c1), f2_add(b.c0, b.c1)), f2_add(t0, t1)), f2_mul_xi(t2));
    r.c2 = f2_add(f2_sub(f2_mul(f2_add(a.c0, a.c2), f2_add(b.c0, b.c2)), f2_add(t0, t2)), t1);
    return r;
}
// a * (b0 + b1 v)
KZG_HD_NOINLINE Fp6 f6_mul_by_01(const Fp6& a, const Fp2& b0, const Fp2& b1) {
    Fp2 t0 = f2_mul(a.c0, b0), t1 = f2_mul(a.c1, b1);
    Fp6 r;
    r.c0 = f2_add(t0, f2_mul_xi(f2_sub(f2_mul(f2_add(a.c1, a.c2), b1), t1)));
    r.c1 = f2_sub(f2_sub(f2_mul(f2_add(a.c0, a.c1), f2_add(b0, b1)), t0), t1);
    r.c2 = f2_add(f2_sub(f2_mul(f2_add(a.c0, a.c2), b0), t0), t1);
    return r;
}
// a * (b1 v)
KZG_HD_NOINLINE Fp6 f6_mul_by_1(const Fp6& a, const Fp2& b1) {
    Fp6 r;
    r.c0 = f2_mul_xi(f2_mul(a.c2, b1));
    r.c1 = f2_mul(a.c0, b1);
    r.c2 = f2_mul(a.c1, b1);
    return r;
}
KZG_HD_NOINLINE Fp6 f6_inv(const Fp6& a) {
    Fp2 c0 = f2_sub(f2_sqr(a.c0), f2_mul_xi(f2_mul(a.c1, a.c2)));
    Fp2 c1 = f2_sub(f2_mul_xi(f2_sqr(a.c2)), f2_mul(a.c0, a.c1));
    Fp2 c2 = f2_sub(f2_sqr(a.c1), f2_mul(a.c0, a.c2));
    Fp2 t = f2_add(f2_mul(a.c0, c0), f2_mul_xi(f2_add(f2_mul(a.c2, c1), f2_mul(a.c1, c2))));
    t = f2_inv(t);
    Fp6 r;
    r.c0 = f2_mul(c0, t);
    r.c1 = f2_mul(c1, t);
    r.c2 = f2_mul(c2, t);
    return r;
}

// ------------------------------------------------------------------------------------------------
// Fp12
// ------------------------------------------------------------------------------------------------
struct Fp12 {
    Fp6 c0, c1;
};
KZG_HD Fp12 f12_one() { Fp12 r; r.c0 = f6_one(); r.c1 = f6_zero(); return r; }
KZG_HD bool f12_is_one(const Fp12& a) { return f6_eq(a.c0, f6_one()) && f6_eq(a.c1, f6_zero()); }
KZG_HD Fp12 f12_conj(const Fp12& a) { Fp12 r; r.c0 = a.c0; r.c1 = f6_neg(a.c1); return r; }

KZG_HD_NOINLINE Fp12 f12_mul(const Fp12& a, const Fp12& b) {
    Fp6 t0 = f6_mul(a.c0, b.c0), t1 = f6_mul(a.c1, b.c1);
    Fp12 r;
    r.c1 = f6_sub(f6_sub(f6_mul(f6_add(a.c0, a.c1), f6_add(b.c0, b.c1)), t0), t1);
    r.c0 = f6_add(t0, f6_mul_v(t1));
    return r;
}
// complex squaring: (a0 + a1 w)^2 = (a0 + a1)(a0 + v a1) - t - v t  +  2t w,  t = a0 a1
KZG_HD_NOINLINE Fp12 f12_sqr(const Fp12& a) {
    Fp6 t = f6_mul(a.c0, a.c1);
    Fp6 s = f6_mul(f6_add(a.c0, a.c1), f6_add(a.c0, f6_mul_v(a.c1)));
    Fp12 r;
    r.c0 = f6_sub(f6_sub(s, t), f6_mul_v(t));
    r.c1 = f6_add(t, t);
    return r;
}
// f * (A + B v + C v w): the sparse product a Miller-loop line needs (13 Fp2 products)
KZG_HD_NOINLINE Fp12 f12_mul_by_line(const Fp12& f, const Fp2& A, const Fp2& B, const Fp2& C) {
    Fp6 t0 = f6_mul_by_01(f.c0, A, B);
    Fp6 t1 = f6_mul_by_1(f.c1, C);
    Fp12 r;
    r.c1 = f6_sub(f6_sub(f6_mul_by_01(f6_add(f.c0, f.c1), A, f2_add(B, C)), t0), t1);
    r.c0 = f6_add(t0, f6_mul_v(t1));
    return r;
}
KZG_HD_NOINLINE Fp12 f12_inv(const Fp12& a) {
    Fp6 t = f6_inv(f6_sub(f6_mul(a.c0, a.c0), f6_mul_v(f6_mul(a.c1, a.c1))));
    Fp12 r;
    r.c0 = f6_mul(a.c0, t);
    r.c1 = f6_neg(f6_mul(a.c1, t));
    return r;
}
// a^p: in the basis sum_k c_k w^k (c_k in Fp2; c0,c2,c4 = a.c0.{c0,c1,c2}, c1,c3,c5 = a.c1.{c0,c1,c2})
// c_k -> conj(c_k) * gamma1[k]
KZG_HD Fp2 frob_gamma1(int k) {
    Fp2 g;
    g.c0 = Fp::from_limbs(FROB_GAMMA1[2 * k]);
    g.c1 = Fp::from_limbs(FROB_GAMMA1[2 * k + 1]);
    return g;
}
KZG_HD_NOINLINE Fp12 f12_frobenius(const Fp12& a) {
    Fp12 r;
    r.c0.c0 = f2_conj(a.c0.c0);
    r.c1.c0 = f2_mul(f2_conj(a.c1.c0), frob_gamma1(1));
    r.c0.c1 = f2_mul(f2_conj(a.c0.c1), frob_gamma1(2));
    r.c1.c1 = f2_mul(f2_conj(a.c1.c1), frob_gamma1(3));
    r.c0.c2 = f2_mul(f2_conj(a.c0.c2), frob_gamma1(4));
    r.c1.c2 = f2_mul(f2_conj(a.c1.c2), frob_gamma1(5));
    return r;
}
// a^(p^2): c_k -> c_k * gamma2[k], gamma2[k] in Fp
KZG_HD_NOINLINE Fp12 f12_frobenius2(const Fp12& a) {
    Fp12 r;
    r.c0.c0 = a.c0.c0;
    r.c1.c0 = f2_mul_fp(a.c1.c0, Fp::from_limbs(FROB_GAMMA2[1]));
    r.c0.c1 = f2_mul_fp(a.c0.c1, Fp::from_limbs(FROB_GAMMA2[2]));
    r.c1.c1 = f2_mul_fp(a.c1.c1, Fp::from_limbs(FROB_GAMMA2[3]));
    r.c0.c2 = f2_mul_fp(a.c0.c2, Fp::from_limbs(FROB_GAMMA2[4]));
    r.c1.c2 = f2_mul_fp(a.c1.c2, Fp::from_limbs(FROB_GAMMA2[5]));
    return r;
}

// Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part of the final
// exponentiation): 9 Fp2 squarings instead of 12 Fp2 products.  Fp4 = Fp2[s]/(s^2 - xi) squaring helper.
KZG_HD void fp4_sqr(Fp2& o0, Fp2& o1, const Fp2& a, const Fp2& b) {
    Fp2 t0 = f2_sqr(a), t1 = f2_sqr(b);
    o0 = f2_add(f2_mul_xi(t1), t0);
    o1 = f2_sub(f2_sub(f2_sqr(f2_add(a, b)), t0), t1);
}
KZG_HD_NOINLINE Fp12 f12_cyclotomic_sqr(const Fp12& f) {
    // coefficients as three Fp4 pairs: (z0,z1) = (c0.c0, c1.c1), (z2,z3) = (c1.c0, c0.c2), (z4,z5) = (c0.c1, c1.c2)
    Fp2 z0 = f.c0.c0, z4 = f.c0.c1, z3 = f.c0.c2, z2 = f.c1.c0, z1 = f.c1.c1, z5 = f.c1.c2;
    Fp2 t0, t1, t2, t3;
    fp4_sqr(t0, t1, z0, z1);
    z0 = f2_add(f2_dbl(f2_sub(t0, z0)), t0);
    z1 = f2_add(f2_dbl(f2_add(t1, z1)), t1);
    fp4_sqr(t0, t1, z2, z3);
    fp4_sqr(t2, t3, z4, z5);
    z4 = f2_add(f2_dbl(f2_sub(t0, z4)), t0);
    z5 = f2_add(f2_dbl(f2_add(t1, z5)), t1);
    t0 = f2_mul_xi(t3);
    z2 = f2_add(f2_dbl(f2_add(t0, z2)), t0);
    z3 = f2_add(f2_dbl(f2_sub(t2, z3)), t2);
    Fp12 r;
    r.c0.c0 = z0; r.c0.c1 = z4; r.c0.c2 = z3;
    r.c1.c0 = z2; r.c1.c1 = z1; r.c1.c2 = z5;
    return r;
}

// ------------------------------------------------------------------------------------------------
// G2 (affine, on the twist y^2 = x^3 + 4(1+u)); setup-time only
// ------------------------------------------------------------------------------------------------
struct G2Affine {
    Fp2 x, y;  // (0,0) = infinity
};
KZG_HD bool g2a_is_inf(const G2Affine& a) { return f2_is_zero(a.x) && f2_is_zero(a.y); }

KZG_HD bool f2_is_lex_largest(const Fp2& y) {
    // sign of y.c1, or of y.c0 when y.c1 == 0 (ZCash G2 encoding)
    if (!is_zero(y.c1)) return fp_is_lex_largest(y.c1);
    return fp_is_lex_largest(y.c0);
}

// blst_p2_uncompress (blst/src/e2.c): 96 bytes = x.c1 || x.c0, flags in byte 0.  No subgroup check
// (load_trusted_setup does none: src/setup/setup.c:469-477).
KZG_HD_NOINLINE bool g2a_uncompress(G2Affine& out, const uint8_t* in) {
    uint8_t b0 = in[0];
    out.x = f2_zero();
    out.y = f2_zero();
    if (!(b0 & 0x80)) return false;
    if (b0 & 0x40) {
        uint32_t acc = b0 & 0x3F;
        for (int i = 1; i < 96; i++) acc |= in[i];
        return acc == 0;
    }
    uint8_t tmp[48];
    for (int i = 0; i < 48; i++) tmp[i] = in[i];
    tmp[0] &= 0x1F;
    uint32_t t1[12], t0[12];
    limbs_from_be<12>(t1, tmp);
    limbs_from_be<12>(t0, in + 48);
    if (limbs_geq<12>(t1, FP_MOD) || limbs_geq<12>(t0, FP_MOD)) return false;
    Fp2 x;
    x.c0 = to_mont<FpTag>(t0);
    x.c1 = to_mont<FpTag>(t1);
    if (f2_is_zero(x)) return false;  // as for G1 (blst reports POINT_NOT_IN_GROUP)
    Fp2 b;
    b.c0 = Fp::from_limbs(FP_B);
    b.c1 = Fp::from_limbs(FP_B);
    Fp2 rhs = f2_add(f2_mul(f2_sqr(x), x), b);
    Fp2 y;
    if (!f2_sqrt(y, rhs)) return false;
    bool want_large = (b0 & 0x20) != 0;
    if (f2_is_lex_largest(y) != want_large) y = f2_neg(y);
    out.x = x;
    out.y = y;
    return true;
}

// ------------------------------------------------------------------------------------------------
// Miller loop with precomputed lines
// ------------------------------------------------------------------------------------------------
constexpr int MILLER_DBL = 63;                       // doublings for |z| = 0xd201000000010000 (64 bits)
constexpr int MILLER_ADD = 5;                        // set bits below the top one
constexpr int MILLER_LINES = MILLER_DBL + MILLER_ADD;

struct LineCoeff {
    Fp2 A, B;  // line(P) = A + (B * xP) v + (yP) v w
};
struct G2Lines {
    LineCoeff line[MILLER_LINES];
    uint32_t is_inf;  // Q at infinity: the pairing is 1
    uint32_t pad[3];
};

// Fill the line table of Q (affine coordinates, one Fp2 inversion per step; runs once per setup).
KZG_HD_NOINLINE void g2_precompute_lines(G2Lines& out, const G2Affine& Q) {
    out.is_inf = g2a_is_inf(Q) ? 1u : 0u;
    out.pad[0] = out.pad[1] = out.pad[2] = 0;
    if (out.is_inf) return;
    Fp2 xT = Q.x, yT = Q.y;
    int k = 0;
    const uint64_t z = BLS_X_ABS;
    for (int b = 62; b >= 0; b--) {
        // tangent at T
        Fp2 x2 = f2_sqr(xT);
        Fp2 lam = f2_mul(f2_add(f2_dbl(x2), x2), f2_inv(f2_dbl(yT)));
        out.line[k].A = f2_sub(f2_mul(lam, xT), yT);
        out.line[k].B = f2_neg(lam);
        k++;
        Fp2 x3 = f2_sub(f2_sqr(lam), f2_dbl(xT));
        yT = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT);
        xT = x3;
        if ((z >> b) & 1ull) {
            // chord through T and Q
            lam = f2_mul(f2_sub(yT, Q.y), f2_inv(f2_sub(xT, Q.x)));
            out.line[k].A = f2_sub(f2_mul(lam, xT), yT);
            out.line[k].B = f2_neg(lam);
            k++;
            x3 = f2_sub(f2_sub(f2_sqr(lam), xT), Q.x);
            yT = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT);
            xT = x3;
        }
    }
}

KZG_HD Fp12 f12_mul_line_at(const Fp12& f, const LineCoeff& l, const G1Affine& P) {
    Fp2 C;
    C.c0 = P.y;
    C.c1 = Fp::zero();
    return f12_mul_by_line(f, l.A, f2_mul_fp(l.B, P.x), C);
}

// prod_i f_{|z|,Q_i}(P_i) for up to two pairs sharing the squarings, conjugated for z < 0.
// A pair whose P or Q is infinity contributes 1.
KZG_HD_NOINLINE Fp12 miller_loop2(const G1Affine& P1, const G2Lines* L1, const G1Affine& P2, const G2Lines* L2) {
    bool use1 = !g1a_is_inf(P1) && !L1->is_inf;
    bool use2 = !g1a_is_inf(P2) && !L2->is_inf;
    Fp12 f = f12_one();
    int k = 0;
    const uint64_t z = BLS_X_ABS;
    for (int b = 62; b >= 0; b--) {
        if (b != 62) f = f12_sqr(f);
        if (use1) f = f12_mul_line_at(f, L1->line[k], P1);
        if (use2) f = f12_mul_line_at(f, L2->line[k], P2);
        k++;
        if ((z >> b) & 1ull) {
            if (use1) f = f12_mul_line_at(f, L1->line[k], P1);
            if (use2) f = f12_mul_line_at(f, L2->line[k], P2);
            k++;
        }
    }
    return f12_conj(f);
}

// f^|z| for f in the cyclotomic subgroup; the caller conjugates for the sign of z
KZG_HD_NOINLINE Fp12 f12_pow_x_abs(const Fp12& f) {
    Fp12 r = f;
    const uint64_t z = BLS_X_ABS;
    for (int b = 62; b >= 0; b--) {
        r = f12_cyclotomic_sqr(r);
        if ((z >> b) & 1ull) r = f12_mul(r, f);
    }
    return r;
}
KZG_HD Fp12 f12_pow_x(const Fp12& f) { return f12_conj(f12_pow_x_abs(f)); }  // z < 0; conj = inverse here

// f^(3 (p^12-1)/r)
KZG_HD_NOINLINE Fp12 final_exponentiation(const Fp12& f_in) {
    // easy part: f^((p^6-1)(p^2+1))
    Fp12 f = f12_mul(f12_conj(f_in), f12_inv(f_in));
    f = f12_mul(f12_frobenius2(f), f);
    // hard part: f^((z-1)^2 (z+p) (z^2+p^2-1) + 3)
    Fp12 t0 = f12_mul(f12_pow_x(f), f12_conj(f));                      // f^(z-1)
    Fp12 t1 = f12_mul(f12_pow_x(t0), f12_conj(t0));                    // f^((z-1)^2)
    Fp12 t2 = f12_mul(f12_pow_x(t1), f12_frobenius(t1));               // ^(z+p)
    Fp12 t3 = f12_mul(f12_mul(f12_pow_x(f12_pow_x(t2)), f12_frobenius2(t2)), f12_conj(t2));  // ^(z^2+p^2-1)
    Fp12 f3 = f12_mul(f12_cyclotomic_sqr(f), f);                       // f^3
    return f12_mul(t3, f3);
}

// e(P1, Q1) * e(P2, Q2) == 1 ?
KZG_HD_NOINLINE bool pairing_product_is_one(const G1Affine& P1, const G2Lines* L1, const G1Affine& P2, const G2Lines* L2) {
    return f12_is_one(final_exponentiation(miller_loop2(P1, L1, P2, L2)));
}

}  // namespace kzg
