// G1 arithmetic for the GPU engine: E1: y^2 = x^3 + 4 over Fp.
//
// Replaces what the reference gets from blst: point formulas (blst/src/ec_ops.h:40-340, XYZZ variants
// :642-787), scalar multiplication (blst/src/ec_mult.h, blst/src/e1.c:396-533), compression
// (blst/src/e1.c:201-294) and the subgroup test (blst/src/map_to_g1.c:512-548), as reached through
// src/common/ec.c:29-58 and src/common/bytes.c:42-116.
//
// Representation: accumulators are extended Jacobian "XYZZ" (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2);
// infinity <=> ZZ == 0.  Table / input points are affine with (0,0) standing for infinity (not on
// the curve, so unambiguous).  Every observable output goes through the canonical 48-byte encoding,
// so internal representation never affects parity.
#pragma once
#include "field.cuh"

namespace kzg {

struct G1Affine {
    Fp x, y;
};
struct G1 {
    Fp x, y, zz, zzz;
};

KZG_HD bool g1a_is_inf(const G1Affine& a) { return is_zero(a.x) && is_zero(a.y); }
KZG_HD G1Affine g1a_inf() {
    G1Affine a;
    a.x = Fp::zero();
    a.y = Fp::zero();
    return a;
}
KZG_HD bool g1_is_inf(const G1& p) { return is_zero(p.zz); }
KZG_HD G1 g1_inf() {
    G1 p;
    p.x = Fp::zero();
    p.y = Fp::one();
    p.zz = Fp::zero();
    p.zzz = Fp::zero();
    return p;
}
KZG_HD G1 g1_from_affine(const G1Affine& a) {
    if (g1a_is_inf(a)) return g1_inf();
    G1 p;
    p.x = a.x;
    p.y = a.y;
    p.zz = Fp::one();
    p.zzz = Fp::one();
    return p;
}
KZG_HD G1 g1_neg(const G1& p) {
    G1 r = p;
    r.y = neg(p.y);
    return r;
}
KZG_HD G1Affine g1a_neg(const G1Affine& a) {
    G1Affine r = a;
    r.y = neg(a.y);
    return r;
}
KZG_HD G1Affine g1a_generator() {
    G1Affine g;
    g.x = Fp::from_limbs(G1_GEN_X);
    g.y = Fp::from_limbs(G1_GEN_Y);
    return g;
}

// 2 * (affine point), result XYZZ.  y == 0 cannot happen on this curve for points of odd order, but
// is handled (-> infinity) for arbitrary curve points.
KZG_HD G1 g1_dbl_affine(const G1Affine& a) {
    if (g1a_is_inf(a) || is_zero(a.y)) return g1_inf();
    Fp U = dbl(a.y);
    Fp V = sqr(U);
    Fp W = mul(U, V);
    Fp S = mul(a.x, V);
    Fp X2 = sqr(a.x);
    Fp M = add(dbl(X2), X2);
    G1 r;
    r.x = sub(sqr(M), dbl(S));
    r.y = sub(mul(M, sub(S, r.x)), mul(W, a.y));
    r.zz = V;
    r.zzz = W;
    return r;
}

// 2 * P  (dbl-2008-s-1, a = 0): 6M + 3S... the cost model in DESIGN.md counts 9 Fp mults
KZG_HD G1 g1_dbl(const G1& p) {
    if (g1_is_inf(p) || is_zero(p.y)) return g1_inf();
    Fp U = dbl(p.y);
    Fp V = sqr(U);
    Fp W = mul(U, V);
    Fp S = mul(p.x, V);
    Fp X2 = sqr(p.x);
    Fp M = add(dbl(X2), X2);
    G1 r;
    r.x = sub(sqr(M), dbl(S));
    r.y = sub(mul(M, sub(S, r.x)), mul(W, p.y));
    r.zz = mul(V, p.zz);
    r.zzz = mul(W, p.zzz);
    return r;
}

// acc += (negate ? -a : a), `a` affine.  Complete: handles acc = inf, a = inf, a = +-acc.
// Hot path of the bucket accumulation: 8M + 2S (madd-2008-s).
KZG_HD void g1_madd(G1& acc, const G1Affine& a_in, bool negate) {
    if (g1a_is_inf(a_in)) return;
    G1Affine a;
    a.x = a_in.x;
    a.y = cneg(a_in.y, negate);
    if (g1_is_inf(acc)) {
        acc.x = a.x;
        acc.y = a.y;
        acc.zz = Fp::one();
        acc.zzz = Fp::one();
        return;
    }
    Fp U2 = mul(a.x, acc.zz);
    Fp S2 = mul(a.y, acc.zzz);
    Fp Pd = sub(U2, acc.x);
    Fp Rd = sub(S2, acc.y);
    if (is_zero(Pd)) {
        if (is_zero(Rd))
            acc = g1_dbl_affine(a);
        else
            acc = g1_inf();
        return;
    }
    Fp PP = sqr(Pd);
    Fp PPP = mul(Pd, PP);
    Fp Q = mul(acc.x, PP);
    Fp X3 = sub(sub(sqr(Rd), PPP), dbl(Q));
    Fp Y3 = sub(mul(Rd, sub(Q, X3)), mul(acc.y, PPP));
    acc.x = X3;
    acc.y = Y3;
    acc.zz = mul(acc.zz, PP);
    acc.zzz = mul(acc.zzz, PPP);
}

// p + q, both XYZZ (add-2008-s): 12M + 2S.  Complete.
KZG_HD G1 g1_add(const G1& p, const G1& q) {
    if (g1_is_inf(p)) return q;
    if (g1_is_inf(q)) return p;
    Fp U1 = mul(p.x, q.zz);
    Fp U2 = mul(q.x, p.zz);
    Fp S1 = mul(p.y, q.zzz);
    Fp S2 = mul(q.y, p.zzz);
    Fp Pd = sub(U2, U1);
    Fp Rd = sub(S2, S1);
    if (is_zero(Pd)) {
        if (is_zero(Rd)) return g1_dbl(p);
        return g1_inf();
    }
    Fp PP = sqr(Pd);
    Fp PPP = mul(Pd, PP);
    Fp Q = mul(U1, PP);
    G1 r;
    r.x = sub(sub(sqr(Rd), PPP), dbl(Q));
    r.y = sub(mul(Rd, sub(Q, r.x)), mul(S1, PPP));
    r.zz = mul(mul(p.zz, q.zz), PP);
    r.zzz = mul(mul(p.zzz, q.zzz), PPP);
    return r;
}

// out-of-line versions for cold code (keeps kernels small)
KZG_HD_NOINLINE void g1_add_to(G1& acc, const G1& q) { acc = g1_add(acc, q); }
KZG_HD_NOINLINE void g1_dbl_to(G1& acc) { acc = g1_dbl(acc); }
KZG_HD_NOINLINE void g1_madd_to(G1& acc, const G1Affine& a, bool negate) { g1_madd(acc, a, negate); }

// XYZZ -> affine (one inversion).  Infinity -> (0,0).
KZG_HD G1Affine g1_to_affine(const G1& p) {
    if (g1_is_inf(p)) return g1a_inf();
    Fp i = fp_inv(mul(p.zz, p.zzz));  // 1/(ZZ*ZZZ)
    G1Affine a;
    a.x = mul(p.x, mul(i, p.zzz));  // X/ZZ
    a.y = mul(p.y, mul(i, p.zz));   // Y/ZZZ
    return a;
}

// projective equality
KZG_HD bool g1_eq(const G1& p, const G1& q) {
    bool pi = g1_is_inf(p), qi = g1_is_inf(q);
    if (pi || qi) return pi && qi;
    return eq(mul(p.x, q.zz), mul(q.x, p.zz)) && eq(mul(p.y, q.zzz), mul(q.y, p.zzz));
}

// [k]P for a plain little-endian scalar of NL limbs (public data; MSB-first double-and-add).
// g1_mul of the reference: src/common/ec.c:53 -> blst_p1_mult (blst/src/e1.c:505).
template <int NL>
KZG_HD_NOINLINE G1 g1_mul_affine(const G1Affine& base, const uint32_t* k) {
    G1 acc = g1_inf();
    for (int i = NL - 1; i >= 0; i--) {
        uint32_t w = k[i];
#pragma unroll 1
        for (int b = 31; b >= 0; b--) {
            g1_dbl_to(acc);
            if ((w >> b) & 1u) g1_madd_to(acc, base, false);
        }
    }
    return acc;
}

// [|z|]P for the curve parameter |z| = 0xd201000000010000
KZG_HD_NOINLINE G1 g1_mul_bls_x(const G1& p) {
    G1 acc = p;
    const uint64_t x = BLS_X_ABS;
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        g1_dbl_to(acc);
        if ((x >> b) & 1ull) g1_add_to(acc, p);
    }
    return acc;
}

KZG_HD bool g1a_on_curve(const G1Affine& a) {
    Fp rhs = add(mul(sqr(a.x), a.x), Fp::from_limbs(FP_B));
    return eq(sqr(a.y), rhs);
}

// Prime-order subgroup membership for a point on the curve (not infinity).
// blst_p1_in_g1 (blst/src/map_to_g1.c:512-548) decides the same predicate with an endomorphism
// trick; here:  P in G1  <=>  phi(P) + [z^2]P == inf  with phi(x,y) = (beta*x, y), beta chosen so
// that phi acts on G1 as -z^2 (M. Scott, "A note on group membership tests for G1, G2 and GT on
// BLS pairing-friendly curves", 2021).  tests/hostcheck pins this against [r]P == inf.
KZG_HD_NOINLINE bool g1a_in_subgroup(const G1Affine& a) {
    G1 p = g1_from_affine(a);
    G1 t = g1_mul_bls_x(g1_mul_bls_x(p));  // [z^2]P  (sign of z cancels)
    G1Affine e;
    e.x = mul(a.x, Fp::from_limbs(FP_BETA_A));
    e.y = a.y;
    g1_madd_to(t, e, false);
    return g1_is_inf(t);
}

// ---- serialisation (ZCash format; blst/src/e1.c:201-294) -------------------------------------
// y "lexicographically largest": plain y > (p-1)/2
KZG_HD bool fp_is_lex_largest(const Fp& y) {
    uint32_t t[12], s[12];
    from_mont<FpTag>(t, y);
    return limbs_sub<12>(s, FP_HALF, t) != 0;  // half - y borrows  <=>  y > half
}

// affine (canonical) -> 48 bytes
KZG_HD void g1a_compress(uint8_t* out, const G1Affine& a) {
    if (g1a_is_inf(a)) {
        out[0] = 0xC0;
        for (int i = 1; i < 48; i++) out[i] = 0;
        return;
    }
    uint32_t t[12];
    from_mont<FpTag>(t, a.x);
    limbs_to_be<12>(out, t);
    out[0] |= 0x80 | (fp_is_lex_largest(a.y) ? 0x20 : 0);
}

// 48 bytes -> affine.  Returns false on any encoding error (blst_p1_uncompress, e1.c:236-294):
// compressed flag missing, malformed infinity, x >= p, x^3+4 not a square.  No subgroup check.
KZG_HD_NOINLINE bool g1a_uncompress(G1Affine& out, const uint8_t* in) {
    uint8_t b0 = in[0];
    out = g1a_inf();
    if (!(b0 & 0x80)) return false;
    if (b0 & 0x40) {
        uint32_t acc = b0 & 0x3F;
        for (int i = 1; i < 48; i++) acc |= in[i];
        return acc == 0;
    }
    uint8_t tmp[48];
    for (int i = 0; i < 48; i++) tmp[i] = in[i];
    tmp[0] &= 0x1F;
    uint32_t t[12];
    limbs_from_be<12>(t, tmp);
    if (limbs_geq<12>(t, FP_MOD)) return false;
    if (limbs_is_zero<12>(t)) return false;  // (0,+-2): blst reports POINT_NOT_IN_GROUP (e1.c:289)
    Fp x = to_mont<FpTag>(t);
    Fp rhs = add(mul(sqr(x), x), Fp::from_limbs(FP_B));
    Fp y = pow_limbs<FpTag, 12>(rhs, FP_SQRT_EXP);
    if (!eq(sqr(y), rhs)) return false;
    bool want_large = (b0 & 0x20) != 0;
    if (fp_is_lex_largest(y) != want_large) y = neg(y);
    out.x = x;
    out.y = y;
    return true;
}

// validate_kzg_g1 (src/common/bytes.c:81): decompress, accept infinity, else require subgroup
KZG_HD bool g1a_validate(G1Affine& out, const uint8_t* in) {
    if (!g1a_uncompress(out, in)) return false;
    if (g1a_is_inf(out)) return true;
    return g1a_in_subgroup(out);
}

// ---- validation that leaves the doubling chains behind (vmsm.cu) --------------------------------
// The subgroup test multiplies by |z| twice.  Done LSB-first, each multiplication walks the pure
// doubling chain D_i = 2^i P and adds D_i for the six set bits of |z| -- the same 63 doublings + 5
// additions as the MSB-first form, but every 2^(8j) multiple of P (first chain) and of Q = [|z|]P
// (second chain) passes by, which is exactly the table the bucket form of the verifiers' linear
// combination needs for the base-|z| expansion of a scalar (k = a0 + a1|z| + a2|z|^2 + a3|z|^3,
// a_i < 2^64; [|z|^2]P = -phi(P), [|z|^3]P = -phi(Q)).  levels[j * stride], j < 9: 2^(8j) P;
// j = 9..17: 2^(8(j-9)) Q.
constexpr int G1_LEVELS = 18;

KZG_HD G1 g1_load(const G1* src) {
#if KZG_DEVICE_PATH
    G1 a;
    const uint4* q = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) d[i] = q[i];
    return a;
#else
    return *src;
#endif
}
KZG_HD void g1_store(G1* dst, const G1& a) {
#if KZG_DEVICE_PATH
    uint4* q = reinterpret_cast<uint4*>(dst);
    const uint4* d = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) q[i] = d[i];
#else
    *dst = a;
#endif
}

KZG_HD_NOINLINE G1 g1_mul_bls_x_levels(const G1& p, G1* levels, size_t stride) {
    G1 d = p, acc = g1_inf();
    const uint64_t x = BLS_X_ABS;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        if ((i & 7) == 0) g1_store(levels + (size_t)(i >> 3) * stride, d);
        if ((x >> i) & 1ull) g1_add_to(acc, d);
        g1_dbl_to(d);
    }
    g1_store(levels + (size_t)8 * stride, d);
    return acc;
}

// g1a_validate + the 18 table levels (all infinity for an invalid or infinite point)
KZG_HD_NOINLINE bool g1a_validate_levels(G1Affine& out, const uint8_t* in, G1* levels, size_t stride) {
    const bool ok = g1a_uncompress(out, in);
    if (!ok || g1a_is_inf(out)) {
        const G1 inf = g1_inf();
#pragma unroll 1
        for (int j = 0; j < G1_LEVELS; j++) g1_store(levels + (size_t)j * stride, inf);
        return ok;
    }
    G1 p = g1_from_affine(out);
    G1 q = g1_mul_bls_x_levels(p, levels, stride);
    G1 t = g1_mul_bls_x_levels(q, levels + (size_t)9 * stride, stride);
    G1Affine e;
    e.x = mul(out.x, Fp::from_limbs(FP_BETA_A));
    e.y = out.y;
    g1_madd_to(t, e, false);  // phi(P) + [z^2]P
    return g1_is_inf(t);
}

// Balanced base-|z| digits of a scalar k < r (8 plain limbs):
//     k = s[0] + s[1]|z| + s[2]|z|^2 + s[3]|z|^3  (mod r),   |s[i]| <= |z|/2 + 1 < 2^62.8.
// Plain digits a[i] < |z| by three long divisions, then a[i] > |z|/2 becomes a[i] - |z| with a carry
// into the next digit; the carry out of the top digit is |z|^4 = |z|^2 - 1 (mod r = z^4 - z^2 + 1).
// Balanced digits never carry out of their top byte when recoded into signed bytes (vmsm.cu).
KZG_HD void basez_split(int64_t s[4], const uint32_t k[8]) {
    const uint64_t Z = BLS_X_ABS, H = BLS_X_ABS / 2;
    uint32_t cur[8], nxt[8];
    uint64_t a[4];
#pragma unroll
    for (int i = 0; i < 8; i++) cur[i] = k[i];
#pragma unroll 1
    for (int step = 0; step < 3; step++) {
        const int nl = 8 - 2 * step;  // the quotient loses 63.7 bits per step
        uint64_t r = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) nxt[i] = 0;
#pragma unroll 1
        for (int bit = 32 * nl - 1; bit >= 0; bit--) {
            const uint64_t top = r >> 63;
            r = (r << 1) | ((cur[bit >> 5] >> (bit & 31)) & 1u);
            if (top || r >= Z) {
                r -= Z;
                nxt[bit >> 5] |= 1u << (bit & 31);
            }
        }
        a[step] = r;
#pragma unroll
        for (int i = 0; i < 8; i++) cur[i] = nxt[i];
    }
    a[3] = ((uint64_t)cur[1] << 32) | cur[0];
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint64_t v = a[i] + carry;  // <= |z|: no overflow
        if (v > H) {
            s[i] = (int64_t)(v - Z);  // two's complement of a value in (-|z|/2, 0]
            carry = 1;
        } else {
            s[i] = (int64_t)v;
            carry = 0;
        }
    }
    s[2] += (int64_t)carry;
    s[0] -= (int64_t)carry;
}

}  // namespace kzg
