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
// A cooperative machine is a GROUP of 64 consecutive threads (two warps) of the CTA, meeting at its own named barrier
// (id 1 + group): a 64-thread CTA is one machine, a 128-thread CTA runs two independent ones side by side (the two
// Miller loops of a pairing check, pairing.cu).
__device__ __forceinline__ void coop_sync() { asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory"); }
#define COOP_BEGIN \
    {              \
        const int lane = threadIdx.x & (COOP_LANES - 1);
#define COOP_END \
    }            \
    coop_sync();
#else
#define COOP_BEGIN for (int lane = 0; lane < COOP_LANES; lane++) {
#define COOP_END }
#endif

// ------------------------------------------------------------------------------------------------
// The "wide" domain.  Inside the cooperative arithmetic an Fp value is any 14-limb integer congruent
// to v * 2^448 (Montgomery radix R_w = 2^448 on 14 limbs), NOT reduced below p.  A tower operation
// spends most of its instructions on the small signed sums around its products (up to 8 + 8 input
// terms per lane, 36 output terms per coefficient); with 67 spare bits those sums are plain
// multi-limb additions -- a negative part is taken off a fixed multiple of p -- and the ONLY modular
// reduction is the one the Montgomery multiplier performs anyway (its result is < p whatever the
// size of the operands, as long as x*y < p * 2^448).  Bounds, in multiples of p (checked against the
// term counts by tools/gen_pairing_tables.py):
//   products and their stored negatives p - v <= 1;  output sums (<= 36 terms) <= 36, stored with their negative
//   64 p - v <= 64 (a conjugation swaps the two copies);  input sums (<= 8 terms) <= 512 < 2^10, so
//   x*y / 2^448 < 2^(2*391-448) << p.
// The previous version reduced after every addition: ~2600 instructions per tower operation, two
// thirds of them in the sums; this one needs ~1200 (14-limb product included).
// Values cross to the 12-limb canonical form (pairing.cuh tower code: Frobenius, the one inversion,
// the final comparison) through one wide product with a constant.
// ------------------------------------------------------------------------------------------------
struct FpWTag {
    static constexpr int N = 14;
    static constexpr uint32_t INV = FP_INV32;
    KZG_HDS const uint32_t* mod() { return FPW_MOD; }
    KZG_HDS const uint32_t* one() { return FPW_ONE; }
    KZG_HDS const uint32_t* r2() { return FPW_R2; }
};
using FpW = Fe<FpWTag>;

KZG_HD FpW w_ext(const Fp& a) {  // the same integer on 14 limbs
    FpW r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = a.l[i];
    r.l[12] = r.l[13] = 0;
    return r;
}
// 12-limb Montgomery form (v * 2^384) -> wide form (v * 2^448): times 2^512 / 2^448
KZG_HD FpW w_from_fp(const Fp& a) { return mul(w_ext(a), FpW::from_limbs(FPW_C512)); }
// wide (any size within the bounds above) -> canonical 12-limb Montgomery form: times 2^384 / 2^448
KZG_HD Fp w_to_fp(const FpW& v) {
    FpW t = mul(v, w_ext(Fp::one()));
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = t.l[i];
    return r;
}
KZG_HD void w_acc(FpW& acc, const FpW& v) { limbs_add<14>(acc.l, acc.l, v.l); }

// Shared-memory form of a wide value: 14 limbs padded to 64 bytes, so that a value is FOUR 128-bit words (4 LDS.128
// instead of 14 LDS.32 per term of a sum -- the sums around the products cost more instructions than the products,
// profiles/R2_summary.md).
struct alignas(16) FpS {
    uint32_t l[16];
};
KZG_HD FpW s_load(const FpS* p) {
    FpW r;
#if KZG_DEVICE_PATH
    const uint4* q = reinterpret_cast<const uint4*>(p->l);
    const uint4 a = q[0], b = q[1], c = q[2], d = q[3];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    r.l[12] = d.x; r.l[13] = d.y;
#else
    for (int i = 0; i < 14; i++) r.l[i] = p->l[i];
#endif
    return r;
}
KZG_HD void s_store(FpS* p, const FpW& v) {
#if KZG_DEVICE_PATH
    uint4* q = reinterpret_cast<uint4*>(p->l);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
    q[3] = make_uint4(v.l[12], v.l[13], 0u, 0u);
#else
    for (int i = 0; i < 14; i++) p->l[i] = v.l[i];
    p->l[14] = p->l[15] = 0;
#endif
}
// the stored negative of a coefficient: 64 p - v (v < 64 p).  Every coefficient and every product is kept in BOTH
// signs, so a signed sum is a plain sum of selected copies: no per-limb sign mask, no offset to start from.
KZG_HD FpW w_neg64(const FpW& v) {
    FpW t;
    limbs_sub<14>(t.l, FPW_OFF64, v.l);
    return t;
}

// lane schedules copied into shared memory at kernel start (a few KB; constant-bank reads through
// generic pointers were the slow part of the first version)
struct CoopTables {
    int16_t off[4][3][56];   // [op][x/y/o][...]
    int8_t xt[4][160], yt[4][160], ot[4][336];
};

struct CoopWS {
    CoopTables tb;
    FpS prod[54], nprod[54];  // the products of the running operation and their negatives (p - v)
    FpS part[12][5];          // partial output sums (5 lanes per output coefficient)
    FpS reg[COOP_NREG][12], nreg[COOP_NREG][12];  // registers and their negatives (64 p - v)
    FpS lv[2][MILLER_LINES][5];  // every line of both pairs evaluated at the (scaled) point: A.c0*s, A.c1*s, B.c0*xs, B.c1*xs, ys
    FpW pt[2][3];    // per pair: s = ZZ*ZZZ and xs = X*ZZZ (times 2^512: see coop_load_points), ys = Y*ZZ (wide)
    FpS onew;        // the constant one, second operand of the cyclotomic square's pass-through products
    FpS zerow;       // padding operand of the sum loops
    Fp canon[12];    // scratch for the excursions into the 12-limb tower code
    int use[2];
    int result;
};
// coefficient `lane` of register d := v (both signs)
KZG_HD void coop_put(CoopWS& ws, int d, int lane, const FpW& v) {
    s_store(&ws.reg[d][lane], v);
    s_store(&ws.nreg[d][lane], w_neg64(v));
}

// One step of a signed sum: code 0 = padding (adds the zero operand), else +-(index + 1): index < 12 selects `a`
// (or, for a negative code, its stored negative `na`), larger indices `b` (never negative: tools/gen_pairing_tables.py
// asserts it).  No branch: every lane of the warp runs the same four loads and the same carry chain.
KZG_HD void coop_acc_term(FpW& acc, int code, const FpS* a, const FpS* na, const FpS* b, const FpS* zero) {
    const int idx = (code < 0 ? -code : code) - 1;
    const FpS* src = (code == 0) ? zero : (idx < 12 ? ((code < 0 ? na : a) + idx) : (b + (idx - 12)));
    const FpW v = s_load(src);
    limbs_add<14>(acc.l, acc.l, v.l);
}
// output sums: the operands are the products (or their negatives)
KZG_HD void coop_acc_prod(FpW& acc, int code, const FpS* prod, const FpS* nprod, const FpS* zero) {
    const int idx = (code < 0 ? -code : code) - 1;
    const FpS* src = (code == 0) ? zero : ((code < 0 ? nprod : prod) + idx);
    const FpW v = s_load(src);
    limbs_add<14>(acc.l, acc.l, v.l);
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
    if (lane == 0) s_store(&ws.onew, FpW::one());
    if (lane == 1) s_store(&ws.zerow, FpW::zero());
    COOP_END
}
KZG_HD CoopOp coop_table(const CoopWS& ws, int op) {
    const int np[4] = {COOP_MUL_NPROD, COOP_SQR_NPROD, COOP_LINE_NPROD, COOP_CYC_NPROD};
    return CoopOp{np[op], ws.tb.off[op][0], ws.tb.off[op][1], ws.tb.off[op][2], ws.tb.xt[op], ws.tb.yt[op], ws.tb.ot[op]};
}

// reg[d] = op(reg[a], b); d may equal a (the outputs read only the products).  `b`: twelve coefficients of another
// register, the five values of a line, or the constant one.  All sums are plain sums of stored copies (see FpS).
KZG_HD_NOINLINE void coop_run(CoopWS& ws, int op, int d, int ra, const FpS* b) {
    const CoopOp T = coop_table(ws, op);
    const FpS *a = ws.reg[ra], *na = ws.nreg[ra];
    COOP_BEGIN
    for (int L = lane; L < T.nprod; L += COOP_LANES) {
        const int bx = T.xo[L], nx = T.xo[L + 1] - bx, by = T.yo[L], ny = T.yo[L + 1] - by;
        const int n = nx > ny ? nx : ny;
        FpW x = FpW::zero(), y = FpW::zero();
        for (int t = 0; t < n; t++) {  // the two chains are independent: they overlap in the pipeline
            coop_acc_term(x, t < nx ? T.xt[bx + t] : 0, a, na, b, &ws.zerow);
            coop_acc_term(y, t < ny ? T.yt[by + t] : 0, a, na, b, &ws.zerow);
        }
        const FpW pr = mul(x, y);  // < p
        s_store(&ws.prod[L], pr);
        FpW npr;
        limbs_sub<14>(npr.l, FPW_MOD, pr.l);  // p - v in (0, p]
        s_store(&ws.nprod[L], npr);
    }
    COOP_END
    // output sums: up to 36 signed products per coefficient -> 5 lanes per coefficient (<= 8 terms each)
    COOP_BEGIN
    if (lane < 60) {
        const int k = lane / 5, sub = lane % 5;
        // two accumulators (alternate terms): two independent carry chains in flight instead of one
        FpW acc = FpW::zero(), acc2 = FpW::zero();
        const int end = T.oo[k + 1];
        for (int t = T.oo[k] + sub; t < end; t += 10) {
            coop_acc_prod(acc, T.ot[t], ws.prod, ws.nprod, &ws.zerow);
            coop_acc_prod(acc2, t + 5 < end ? T.ot[t + 5] : 0, ws.prod, ws.nprod, &ws.zerow);
        }
        w_acc(acc, acc2);
        s_store(&ws.part[k][sub], acc);  // <= 8 p
    }
    COOP_END
    COOP_BEGIN
    if (lane < 12) {
        FpW u = s_load(&ws.part[lane][0]), v = s_load(&ws.part[lane][2]);
        w_acc(u, s_load(&ws.part[lane][1]));
        w_acc(v, s_load(&ws.part[lane][3]));
        w_acc(u, s_load(&ws.part[lane][4]));
        w_acc(u, v);
        coop_put(ws, d, lane, u);  // <= 36 p, and 64 p - u
    }
    COOP_END
}

KZG_HD void coop_mul(CoopWS& ws, int d, int a, int b) { coop_run(ws, COOP_OP_MUL, d, a, ws.reg[b]); }
KZG_HD void coop_sqr(CoopWS& ws, int d, int a) { coop_run(ws, COOP_OP_SQR, d, a, ws.reg[a]); }
KZG_HD void coop_cyc(CoopWS& ws, int d, int a) { coop_run(ws, COOP_OP_CYC, d, a, &ws.onew); }
KZG_HD void coop_line(CoopWS& ws, int d, int a, int pair, int k) { coop_run(ws, COOP_OP_LINE, d, a, ws.lv[pair][k]); }

// conjugation over Fp6: negate the coefficients of the odd powers of w = swap their two stored copies
KZG_HD_NOINLINE void coop_conj(CoopWS& ws, int d, int a) {
    COOP_BEGIN
    if (lane < 12) {
        const int k = lane >> 1;
        const FpW v = s_load(&ws.reg[a][lane]), nv = s_load(&ws.nreg[a][lane]);
        s_store(&ws.reg[d][lane], (k & 1) ? nv : v);
        s_store(&ws.nreg[d][lane], (k & 1) ? v : nv);
    }
    COOP_END
}
KZG_HD_NOINLINE void coop_copy(CoopWS& ws, int d, int a) {
    COOP_BEGIN
    if (lane < 12) {
        const FpW v = s_load(&ws.reg[a][lane]), nv = s_load(&ws.nreg[a][lane]);
        s_store(&ws.reg[d][lane], v);
        s_store(&ws.nreg[d][lane], nv);
    }
    COOP_END
}
KZG_HD void coop_set_one(CoopWS& ws, int d) {
    COOP_BEGIN
    if (lane < 12) coop_put(ws, d, lane, (lane == 0) ? FpW::one() : FpW::zero());
    COOP_END
}
// a^(p^power), power = 1 or 2; d != a.  Lane k < 6 owns the Fp2 coefficient of w^k (12-limb tower code).
KZG_HD_NOINLINE void coop_frobenius(CoopWS& ws, int d, int a, int power) {
    COOP_BEGIN
    if (lane < 6) {
        Fp2 c;
        c.c0 = w_to_fp(s_load(&ws.reg[a][2 * lane]));
        c.c1 = w_to_fp(s_load(&ws.reg[a][2 * lane + 1]));
        if (power == 1) {
            c = f2_conj(c);
            if (lane != 0) c = f2_mul(c, frob_gamma1(lane));
        } else if (lane != 0) {
            c = f2_mul_fp(c, Fp::from_limbs(FROB_GAMMA2[lane]));
        }
        coop_put(ws, d, 2 * lane, w_from_fp(c.c0));
        coop_put(ws, d, 2 * lane + 1, w_from_fp(c.c1));
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
// the one inversion of the final exponentiation: serial on lane 0, in the 12-limb tower code
KZG_HD_NOINLINE void coop_inv(CoopWS& ws, int d, int a) {
    COOP_BEGIN
    if (lane < 12) ws.canon[lane] = w_to_fp(s_load(&ws.reg[a][lane]));
    COOP_END
    COOP_BEGIN
    if (lane == 0) {
        Fp12 v = coop_to_tower(ws.canon);
        coop_from_tower(ws.canon, f12_inv(v));
    }
    COOP_END
    COOP_BEGIN
    if (lane < 12) coop_put(ws, d, lane, w_from_fp(ws.canon[lane]));
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
        // s and xs meet a 12-limb line coefficient (c * 2^384) in one WIDE product that must come out as
        // c*s * 2^448: carry them as s * 2^512, i.e. times 2^576 / 2^448.  ys enters the line as it is.
        ws.pt[pair][j] = (j == 2) ? w_from_fp(v) : mul(w_ext(v), FpW::from_limbs(FPW_C576));
    }
    if (lane == 6) ws.use[0] = (!g1_is_inf(P1) && !L1->is_inf) ? 1 : 0;
    if (lane == 7) ws.use[1] = (!g1_is_inf(P2) && !L2->is_inf) ? 1 : 0;
    COOP_END
}

// Every line of both pairs evaluated at the (scaled) points -> ws.lv, in one parallel pass before the
// Miller loop (544 independent products over the 64 lanes).  The first version did this step by step
// inside the loop: 68 extra barriers and 68 exposed global-memory round trips, 15 % of the kernel.
KZG_HD_NOINLINE void coop_prepare_all_lines(CoopWS& ws, const G2Lines* L1, const G2Lines* L2) {
    COOP_BEGIN
    for (int i = lane; i < 2 * MILLER_LINES * 5; i += COOP_LANES) {
        const int pair = i / (MILLER_LINES * 5), r = i % (MILLER_LINES * 5), k = r / 5, j = r % 5;
        if (!ws.use[pair]) continue;
        const LineCoeff& l = (pair ? L2 : L1)->line[k];
        FpW v;
        if (j == 0) v = mul(w_ext(l.A.c0), ws.pt[pair][0]);
        else if (j == 1) v = mul(w_ext(l.A.c1), ws.pt[pair][0]);
        else if (j == 2) v = mul(w_ext(l.B.c0), ws.pt[pair][1]);
        else if (j == 3) v = mul(w_ext(l.B.c1), ws.pt[pair][1]);
        else v = ws.pt[pair][2];
        s_store(&ws.lv[pair][k][j], v);
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

// reg[F] = conj(Miller loop) over the pairs whose ws.use[] flag is set (lines already evaluated: coop_prepare_all_lines)
KZG_HD void coop_miller_loop(CoopWS& ws) {
    enum { F = 0 };
    const bool use0 = ws.use[0] != 0, use1 = ws.use[1] != 0;
    coop_set_one(ws, F);
    const uint64_t z = BLS_X_ABS;
    int k = 0;
    for (int b = 62; b >= 0; b--) {
        if (b != 62) coop_sqr(ws, F, F);
        if (use0) coop_line(ws, F, F, 0, k);
        if (use1) coop_line(ws, F, F, 1, k);
        k++;
        if ((z >> b) & 1ull) {
            if (use0) coop_line(ws, F, F, 0, k);
            if (use1) coop_line(ws, F, F, 1, k);
            k++;
        }
    }
    coop_conj(ws, F, F);
}
KZG_HD void coop_final_exp_is_one(CoopWS& ws);

// ws.result = [ e(+-P1, Q1) * e(P2, Q2) == 1 ]
KZG_HD void coop_pairing_product_is_one(CoopWS& ws, const G1& P1, const G2Lines* L1, const G1& P2, const G2Lines* L2, bool negate_first) {
    coop_init_tables(ws);
    coop_load_points(ws, P1, L1, P2, L2, negate_first);
    coop_prepare_all_lines(ws, L1, L2);
    coop_miller_loop(ws);
    coop_final_exp_is_one(ws);
}

// ws.result = [ reg[F]^((p^12 - 1) / r) == 1 ]
KZG_HD void coop_final_exp_is_one(CoopWS& ws) {
    enum { F = 0, E = 1, T0 = 2, T1 = 3, T2 = 4, T3 = 5, X = 6, Y = 7 };
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
    if (lane < 12) ws.canon[lane] = w_to_fp(s_load(&ws.reg[T3][lane]));
    COOP_END
    COOP_BEGIN
    if (lane == 0) {
        bool one = eq(ws.canon[0], Fp::one());
        for (int i = 1; i < 12; i++) one = one && is_zero(ws.canon[i]);
        ws.result = one ? 1 : 0;
    }
    COOP_END
}

}  // namespace kzg
