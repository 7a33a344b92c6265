// Thread-cooperative pairing check: the Miller loop (precomputed lines) and the final exponentiation
// of pairing.cuh, re-scheduled so that the ~54 base-field products inside every Fp12 operation run
// on separate lanes of one 64-thread CTA instead of back to back on one thread.
//
// Why: a verify call contains exactly one pairing check (src/common/utils.c:172-196), ~19k DEPENDENT
// base-field products -- 22.8 ms on a single GPU thread (profiles/r01), 33 % of a 4096-blob
// verification and ~90 % of a 64-blob one.  Each tower operation (Fp12 product, square, sparse line
// product, cyclotomic square) is bilinear of depth one, so tools/gen_pairing_tables.py flattens it
// into a lane schedule (pairing_tables.cuh): phase 1, lane L forms two small signed sums of input
// coefficients and multiplies them; phase 2, twelve lanes gather the output coefficients.  Operands
// live in shared memory as 12 Fp coefficients in w-power order (index 2k+j).
//
// Inversion-free G1 handling: the points arrive as XYZZ sums; a line A + B*x v + y v w evaluated at
// (X/ZZ, Y/ZZZ) is scaled by ZZ*ZZZ (an Fp factor, killed by the final exponentiation) into
// A*(ZZ ZZZ) + B*(X ZZZ) v + (Y ZZ) v w, so no to-affine inversion is needed.
//
// The same source compiles for the host (tests/hostcheck): COOP blocks become loops over the lanes,
// which is how the schedule is validated in the GPU-less build container.
#pragma once
#include "pairing.cuh"
#include "pairing_tables.cuh"

namespace kzg {

constexpr int COOP_LANES = 64;
constexpr int COOP_NREG = 8;

#if KZG_DEVICE_PATH
#define COOP_BEGIN \
    {              \
        const int lane = threadIdx.x;
#define COOP_END      \
    }                 \
    __syncthreads();
#else
#define COOP_BEGIN for (int lane = 0; lane < COOP_LANES; lane++) {
#define COOP_END }
#endif

// lane schedules copied into shared memory at kernel start (a few KB; constant-bank reads through
// generic pointers were the slow part of the first version)
struct CoopTables {
    int16_t off[4][3][56];   // [op][x/y/o][...]
    int8_t xt[4][160], yt[4][160], ot[4][336];
};

struct CoopWS {
    CoopTables tb;
    Fp prod[54];
    Fp part[12][5];  // partial output sums (5 lanes per output coefficient)
    Fp reg[COOP_NREG][12];
    Fp line[2][5];  // per pair: A.c0*s, A.c1*s, B.c0*xs, B.c1*xs, ys
    Fp pt[2][3];    // per pair: s = ZZ*ZZZ, xs = X*ZZZ, ys = Y*ZZ
    int use[2];
    int result;
};

KZG_HD Fp sub_(const Fp& x, const Fp& y) { return sub(x, y); }

// signed sum of inputs (indices 0..11 -> a, 12.. -> b)
KZG_HD Fp coop_sum_inputs(const int8_t* terms, int beg, int end, const Fp* a, const Fp* b) {
    Fp acc = Fp::zero();
    for (int t = beg; t < end; t++) {
        int code = terms[t];
        int idx = (code < 0 ? -code : code) - 1;
        const Fp& v = (idx < 12) ? a[idx] : b[idx - 12];
        acc = (code > 0) ? add(acc, v) : sub(acc, v);
    }
    return acc;
}

struct CoopOp {
    int nprod;
    const int16_t *xo, *yo, *oo;
    const int8_t *xt, *yt, *ot;
};
enum { COOP_OP_MUL = 0, COOP_OP_SQR = 1, COOP_OP_LINE = 2, COOP_OP_CYC = 3 };

KZG_HD void coop_copy_table(CoopTables& tb, int op, int lane, int nprod, const int16_t* xo, const int16_t* yo, const int16_t* oo, const int8_t* xt, const int8_t* yt,
                            const int8_t* ot) {
    for (int i = lane; i <= nprod; i += COOP_LANES) {
        tb.off[op][0][i] = xo[i];
        tb.off[op][1][i] = yo[i];
    }
    for (int i = lane; i <= 12; i += COOP_LANES) tb.off[op][2][i] = oo[i];
    for (int i = lane; i < xo[nprod]; i += COOP_LANES) tb.xt[op][i] = xt[i];
    for (int i = lane; i < yo[nprod]; i += COOP_LANES) tb.yt[op][i] = yt[i];
    for (int i = lane; i < oo[12]; i += COOP_LANES) tb.ot[op][i] = ot[i];
}
// must run once before any coop_run
KZG_HD void coop_init_tables(CoopWS& ws) {
    COOP_BEGIN
    coop_copy_table(ws.tb, COOP_OP_MUL, lane, COOP_MUL_NPROD, COOP_MUL_XOFF, COOP_MUL_YOFF, COOP_MUL_OOFF, COOP_MUL_XT, COOP_MUL_YT, COOP_MUL_OT);
    coop_copy_table(ws.tb, COOP_OP_SQR, lane, COOP_SQR_NPROD, COOP_SQR_XOFF, COOP_SQR_YOFF, COOP_SQR_OOFF, COOP_SQR_XT, COOP_SQR_YT, COOP_SQR_OT);
    coop_copy_table(ws.tb, COOP_OP_LINE, lane, COOP_LINE_NPROD, COOP_LINE_XOFF, COOP_LINE_YOFF, COOP_LINE_OOFF, COOP_LINE_XT, COOP_LINE_YT, COOP_LINE_OT);
    coop_copy_table(ws.tb, COOP_OP_CYC, lane, COOP_CYC_NPROD, COOP_CYC_XOFF, COOP_CYC_YOFF, COOP_CYC_OOFF, COOP_CYC_XT, COOP_CYC_YT, COOP_CYC_OT);
    COOP_END
}
KZG_HD CoopOp coop_table(const CoopWS& ws, int op) {
    const int np[4] = {COOP_MUL_NPROD, COOP_SQR_NPROD, COOP_LINE_NPROD, COOP_CYC_NPROD};
    return CoopOp{np[op], ws.tb.off[op][0], ws.tb.off[op][1], ws.tb.off[op][2], ws.tb.xt[op], ws.tb.yt[op], ws.tb.ot[op]};
}

// dst = op(a, b); dst may alias a or b (outputs only read the products, and -- cyclotomic square --
// the SAME coefficient of a that the lane overwrites).
KZG_HD_NOINLINE void coop_run(CoopWS& ws, int op, Fp* dst, const Fp* a, const Fp* b) {
    const CoopOp T = coop_table(ws, op);
    COOP_BEGIN
    for (int L = lane; L < T.nprod; L += COOP_LANES) {
        Fp x = coop_sum_inputs(T.xt, T.xo[L], T.xo[L + 1], a, b);
        Fp y = coop_sum_inputs(T.yt, T.yo[L], T.yo[L + 1], a, b);
        ws.prod[L] = mul(x, y);
    }
    COOP_END
    // output sums: up to 36 signed terms per coefficient -> 5 lanes per coefficient, then a 5-term tail
    COOP_BEGIN
    if (lane < 60) {
        const int k = lane / 5, sub = lane % 5;
        Fp acc = Fp::zero();
        for (int t = T.oo[k] + sub; t < T.oo[k + 1]; t += 5) {
            int code = T.ot[t];
            int mag = code < 0 ? -code : code;
            const Fp& v = (mag > 64) ? a[mag - 65] : ws.prod[mag - 1];
            acc = (code > 0) ? add(acc, v) : sub_(acc, v);
        }
        ws.part[k][sub] = acc;
    }
    COOP_END
    COOP_BEGIN
    if (lane < 12) {
        Fp acc = add(add(ws.part[lane][0], ws.part[lane][1]), add(ws.part[lane][2], ws.part[lane][3]));
        dst[lane] = add(acc, ws.part[lane][4]);
    }
    COOP_END
}

KZG_HD void coop_mul(CoopWS& ws, int d, int a, int b) { coop_run(ws, COOP_OP_MUL, ws.reg[d], ws.reg[a], ws.reg[b]); }
KZG_HD void coop_sqr(CoopWS& ws, int d, int a) { coop_run(ws, COOP_OP_SQR, ws.reg[d], ws.reg[a], ws.reg[a]); }
KZG_HD void coop_cyc(CoopWS& ws, int d, int a) { coop_run(ws, COOP_OP_CYC, ws.reg[d], ws.reg[a], ws.reg[a]); }
KZG_HD void coop_line(CoopWS& ws, int d, int a, int pair) { coop_run(ws, COOP_OP_LINE, ws.reg[d], ws.reg[a], ws.line[pair]); }

// conjugation over Fp6: negate the coefficients of the odd powers of w
KZG_HD_NOINLINE void coop_conj(CoopWS& ws, int d, int a) {
    COOP_BEGIN
    if (lane < 12) {
        int k = lane >> 1;
        ws.reg[d][lane] = (k & 1) ? neg(ws.reg[a][lane]) : ws.reg[a][lane];
    }
    COOP_END
}
KZG_HD_NOINLINE void coop_copy(CoopWS& ws, int d, int a) {
    COOP_BEGIN
    if (lane < 12) ws.reg[d][lane] = ws.reg[a][lane];
    COOP_END
}
KZG_HD void coop_set_one(CoopWS& ws, int d) {
    COOP_BEGIN
    if (lane < 12) ws.reg[d][lane] = (lane == 0) ? Fp::one() : Fp::zero();
    COOP_END
}
// a^(p^power), power = 1 or 2; d != a.  Lane k < 6 owns the Fp2 coefficient of w^k.
KZG_HD_NOINLINE void coop_frobenius(CoopWS& ws, int d, int a, int power) {
    COOP_BEGIN
    if (lane < 6) {
        Fp2 c;
        c.c0 = ws.reg[a][2 * lane];
        c.c1 = ws.reg[a][2 * lane + 1];
        if (power == 1) {
            c = f2_conj(c);
            if (lane != 0) c = f2_mul(c, frob_gamma1(lane));
        } else if (lane != 0) {
            c = f2_mul_fp(c, Fp::from_limbs(FROB_GAMMA2[lane]));
        }
        ws.reg[d][2 * lane] = c.c0;
        ws.reg[d][2 * lane + 1] = c.c1;
    }
    COOP_END
}

// flat (w-power order) <-> tower struct
KZG_HD Fp12 coop_to_tower(const Fp* c) {
    Fp12 r;
    r.c0.c0.c0 = c[0]; r.c0.c0.c1 = c[1];
    r.c1.c0.c0 = c[2]; r.c1.c0.c1 = c[3];
    r.c0.c1.c0 = c[4]; r.c0.c1.c1 = c[5];
    r.c1.c1.c0 = c[6]; r.c1.c1.c1 = c[7];
    r.c0.c2.c0 = c[8]; r.c0.c2.c1 = c[9];
    r.c1.c2.c0 = c[10]; r.c1.c2.c1 = c[11];
    return r;
}
KZG_HD void coop_from_tower(Fp* c, const Fp12& r) {
    c[0] = r.c0.c0.c0; c[1] = r.c0.c0.c1;
    c[2] = r.c1.c0.c0; c[3] = r.c1.c0.c1;
    c[4] = r.c0.c1.c0; c[5] = r.c0.c1.c1;
    c[6] = r.c1.c1.c0; c[7] = r.c1.c1.c1;
    c[8] = r.c0.c2.c0; c[9] = r.c0.c2.c1;
    c[10] = r.c1.c2.c0; c[11] = r.c1.c2.c1;
}
// the one inversion of the final exponentiation: serial on lane 0
KZG_HD_NOINLINE void coop_inv(CoopWS& ws, int d, int a) {
    COOP_BEGIN
    if (lane == 0) {
        Fp12 v = coop_to_tower(ws.reg[a]);
        coop_from_tower(ws.reg[d], f12_inv(v));
    }
    COOP_END
}

// Load the two G1 arguments (XYZZ).  `negate_first`: use -P1 (the e(-a1,a2) of pairings_verify).
KZG_HD void coop_load_points(CoopWS& ws, const G1& P1, const G2Lines* L1, const G1& P2, const G2Lines* L2, bool negate_first) {
    COOP_BEGIN
    if (lane < 6) {
        int pair = lane / 3, j = lane % 3;
        const G1& Pt = pair ? P2 : P1;
        Fp v;
        if (j == 0) v = mul(Pt.zz, Pt.zzz);
        else if (j == 1) v = mul(Pt.x, Pt.zzz);
        else {
            v = mul(Pt.y, Pt.zz);
            if (pair == 0 && negate_first) v = neg(v);
        }
        ws.pt[pair][j] = v;
    }
    if (lane == 6) ws.use[0] = (!g1_is_inf(P1) && !L1->is_inf) ? 1 : 0;
    if (lane == 7) ws.use[1] = (!g1_is_inf(P2) && !L2->is_inf) ? 1 : 0;
    COOP_END
}

// line k of both pairs, evaluated at the (scaled) points -> ws.line
KZG_HD_NOINLINE void coop_prepare_lines(CoopWS& ws, const G2Lines* L1, const G2Lines* L2, int k) {
    COOP_BEGIN
    if (lane < 10) {
        int pair = lane / 5, j = lane % 5;
        const LineCoeff& l = (pair ? L2 : L1)->line[k];
        Fp v;
        if (j == 0) v = mul(l.A.c0, ws.pt[pair][0]);
        else if (j == 1) v = mul(l.A.c1, ws.pt[pair][0]);
        else if (j == 2) v = mul(l.B.c0, ws.pt[pair][1]);
        else if (j == 3) v = mul(l.B.c1, ws.pt[pair][1]);
        else v = ws.pt[pair][2];
        ws.line[pair][j] = v;
    }
    COOP_END
}

// reg[d] = reg[s]^z for the (negative) curve parameter; d != s; reg[s] in the cyclotomic subgroup
KZG_HD_NOINLINE void coop_pow_x(CoopWS& ws, int d, int s) {
    coop_copy(ws, d, s);
    const uint64_t z = BLS_X_ABS;
    for (int b = 62; b >= 0; b--) {
        coop_cyc(ws, d, d);
        if ((z >> b) & 1ull) coop_mul(ws, d, d, s);
    }
    coop_conj(ws, d, d);
}

// ws.result = [ e(+-P1, Q1) * e(P2, Q2) == 1 ]
KZG_HD void coop_pairing_product_is_one(CoopWS& ws, const G1& P1, const G2Lines* L1, const G1& P2, const G2Lines* L2, bool negate_first) {
    enum { F = 0, E = 1, T0 = 2, T1 = 3, T2 = 4, T3 = 5, X = 6, Y = 7 };
    coop_init_tables(ws);
    coop_load_points(ws, P1, L1, P2, L2, negate_first);
    const bool use0 = ws.use[0] != 0, use1 = ws.use[1] != 0;
    // ---- Miller loop ----
    coop_set_one(ws, F);
    const uint64_t z = BLS_X_ABS;
    int k = 0;
    for (int b = 62; b >= 0; b--) {
        if (b != 62) coop_sqr(ws, F, F);
        coop_prepare_lines(ws, L1, L2, k);
        if (use0) coop_line(ws, F, F, 0);
        if (use1) coop_line(ws, F, F, 1);
        k++;
        if ((z >> b) & 1ull) {
            coop_prepare_lines(ws, L1, L2, k);
            if (use0) coop_line(ws, F, F, 0);
            if (use1) coop_line(ws, F, F, 1);
            k++;
        }
    }
    coop_conj(ws, F, F);
    // ---- final exponentiation, easy part: E = F^((p^6-1)(p^2+1)) ----
    coop_inv(ws, X, F);
    coop_conj(ws, Y, F);
    coop_mul(ws, E, Y, X);
    coop_frobenius(ws, X, E, 2);
    coop_mul(ws, E, X, E);
    // ---- hard part: E^((z-1)^2 (z+p)(z^2+p^2-1) + 3) ----
    coop_pow_x(ws, T0, E);            // E^z
    coop_conj(ws, X, E);
    coop_mul(ws, T0, T0, X);          // E^(z-1)
    coop_pow_x(ws, T1, T0);
    coop_conj(ws, X, T0);
    coop_mul(ws, T1, T1, X);          // ^(z-1)^2
    coop_pow_x(ws, T2, T1);
    coop_frobenius(ws, X, T1, 1);
    coop_mul(ws, T2, T2, X);          // ^(z+p)
    coop_pow_x(ws, X, T2);
    coop_pow_x(ws, T3, X);            // T2^(z^2)
    coop_frobenius(ws, X, T2, 2);
    coop_mul(ws, T3, T3, X);
    coop_conj(ws, X, T2);
    coop_mul(ws, T3, T3, X);          // ^(z^2+p^2-1)
    coop_cyc(ws, X, E);
    coop_mul(ws, X, X, E);            // E^3
    coop_mul(ws, T3, T3, X);
    COOP_BEGIN
    if (lane == 0) {
        bool one = eq(ws.reg[T3][0], Fp::one());
        for (int i = 1; i < 12; i++) one = one && is_zero(ws.reg[T3][i]);
        ws.result = one ? 1 : 0;
    }
    COOP_END
}

}  // namespace kzg
