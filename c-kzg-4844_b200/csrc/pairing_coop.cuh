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
inline void coop_sync() {}
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

// ------------------------------------------------------------------------------------------------
// The wide product in radix 2^29 (compile with -DCOOP_W29=1; MEASURED SLOWER, off).  The 32-bit-limb Montgomery product
// of field.cuh is a sequence of carry chains -- every multiply-add waits for the carry of the one before, PTX has a
// single carry flag, and one warp alone on its scheduler reaches 65 % of the multiplier's rate (1.25 us per 14-limb
// product, half of a cyclotomic square).  With 29-bit limbs a column of 14 products (< 2^58 each) and the 14 reduction
// terms under it fit a 64-bit accumulator, so the 392 multiply-adds are INDEPENDENT IMAD.WIDE instructions with 64-bit
// accumulation, no carry flag anywhere; the carries are resolved once at the end (split every column in three 29-bit
// pieces, one short 32-bit carry chain, repack to 32-bit words).  R_w = 2^(29 * 14) = 2^406: for operands < 2^391 the
// result is < p + 2^376 -- NOT reduced below p; the stored negative of a product is then 2 p - v
// (tools/gen_pairing_tables.py checks the totals).  Plain C++: the same code runs in tests/hostcheck.
// Measured on B200 (profiles/pairing_probe_R3e.log): cyclotomic square 2.63 us against 2.36 us, product 3.81 against
// 3.47, whole check 1.41 ms against 1.26 ms.  The 392 wide multiply-adds alone are 1570 cycles of the quarter-rate
// multiplier, and splitting / recombining the limbs adds ~370 instructions per product (1488 against 1112 per
// cyclotomic square): what the carry chains lose in latency the conversions lose in issue slots.  Kept, off, with its
// host test, as the record of the experiment.
// ------------------------------------------------------------------------------------------------
#ifndef COOP_W29
#define COOP_W29 0
#endif
constexpr uint32_t W29_MASK = (1u << 29) - 1;
// 14 radix-2^32 words (value < 2^406) -> 14 radix-2^29 limbs
KZG_HD void w_split29(uint32_t* o, const uint32_t* l) {
#pragma unroll
    for (int k = 0; k < 14; k++) {
        const int bit = 29 * k, w = bit >> 5, sh = bit & 31;
        uint32_t v = l[w] >> sh;
        if (sh > 3 && w + 1 < 14) v |= l[w + 1] << (32 - sh);
        o[k] = v & W29_MASK;
    }
}
KZG_HD void w_mul29(uint32_t* out, const uint32_t* x, const uint32_t* y) {
    uint32_t a[14], b[14];
    w_split29(a, x);
    w_split29(b, y);
    uint64_t t[28];
#pragma unroll
    for (int k = 0; k < 28; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 14; i++)
#pragma unroll
        for (int j = 0; j < 14; j++) t[i + j] += (uint64_t)a[i] * b[j];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
        const uint64_t s = t[i] + c;
        const uint32_t m = ((uint32_t)s * FPW_INV29) & W29_MASK;
        c = (s + (uint64_t)m * FPW_P29[0]) >> 29;  // the low 29 bits cancel
#pragma unroll
        for (int j = 1; j < 14; j++) t[i + j] += (uint64_t)m * FPW_P29[j];
    }
    t[14] += c;
    // columns 14..27 (< 2^63 each) -> lazy 29-bit limbs (three pieces per column) -> exact limbs (one carry chain)
    uint32_t q[17];
#pragma unroll
    for (int k = 0; k < 17; k++) q[k] = 0;
#pragma unroll
    for (int k = 0; k < 14; k++) {
        const uint64_t v = t[14 + k];
        q[k] += (uint32_t)v & W29_MASK;
        q[k + 1] += (uint32_t)(v >> 29) & W29_MASK;
        q[k + 2] += (uint32_t)(v >> 58);
    }
    uint32_t r[17], cc = 0;
#pragma unroll
    for (int k = 0; k < 17; k++) {
        const uint32_t sum = q[k] + cc;
        r[k] = sum & W29_MASK;
        cc = sum >> 29;
    }
    // 29-bit limbs -> 32-bit words (the value is < 2 p < 2^382: twelve words, the top two are zero)
#pragma unroll
    for (int w = 0; w < 14; w++) {
        const int bit = 32 * w, k = bit / 29, off = bit - 29 * k;
        uint64_t v = (uint64_t)r[k] >> off;
        if (k + 1 < 17) v |= (uint64_t)r[k + 1] << (29 - off);
        if (k + 2 < 17) v |= (uint64_t)r[k + 2] << (58 - off);
        out[w] = (uint32_t)v;
    }
}

KZG_HD FpW w_ext(const Fp& a) {  // the same integer on 14 limbs
    FpW r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = a.l[i];
    r.l[12] = r.l[13] = 0;
    return r;
}
// the wide product: x * y / R_w mod p for operands within the bounds above
KZG_HD FpW w_mul(const FpW& x, const FpW& y) {
#if COOP_W29
    FpW r;
    w_mul29(r.l, x.l, y.l);  // < p + 2^376
    return r;
#else
    return mul(x, y);  // < p
#endif
}
#if COOP_W29
#define FPW_CONST_ONE FPW29_ONE
#define FPW_CONST_CIN FPW29_CIN
#define FPW_CONST_CPT FPW29_CPT
#define FPW_CONST_PNEG FPW_2P /* stored negative of a product: 2 p - v */
#else
#define FPW_CONST_ONE FPW_ONE
#define FPW_CONST_CIN FPW_C512
#define FPW_CONST_CPT FPW_C576
#define FPW_CONST_PNEG FPW_MOD /* products are < p: p - v */
#endif
KZG_HD FpW w_one() { return FpW::from_limbs(FPW_CONST_ONE); }
// 12-limb Montgomery form (v * 2^384) -> wide form (v * R_w): one wide product with 2^(2 RW - 384)
KZG_HD FpW w_from_fp(const Fp& a) { return w_mul(w_ext(a), FpW::from_limbs(FPW_CONST_CIN)); }
// wide (any size within the bounds above) -> canonical 12-limb Montgomery form: times 2^384 / R_w
KZG_HD Fp w_to_fp(const FpW& v) {
    FpW t = w_mul(v, w_ext(Fp::one()));
#if COOP_W29
    FpW u;
    if (limbs_sub<14>(u.l, t.l, FPW_MOD) == 0) t = u;  // < 2 p -> < p
#endif
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = t.l[i];
    return r;
}
KZG_HD void w_acc(FpW& acc, const FpW& v) { limbs_add<14>(acc.l, acc.l, v.l); }

// Shared-memory form of a wide value: 14 limbs padded to 64 bytes, so that a value is FOUR 128-bit words (4 LDS.128
// instead of 14 LDS.32 per term of a sum -- the sums around the products cost more instructions than the products,
// profiles/R2_summary.md).
// Stride: 80 bytes, not 64.  A 128-bit shared-memory load is served a quarter warp at a time, and eight lanes reading
// the same quarter of eight different values 64 bytes apart fall into two groups of four banks: four-way conflicts on
// nearly every load of a sum (ncu R2u: 4.5 wavefronts per load instruction, 10.7 per store).  With 80 bytes the bank
// group is (5 e + q) mod 8 for value e: eight consecutive values never collide, arbitrary ones rarely.
#ifndef COOP_FPS_WORDS
#define COOP_FPS_WORDS 20
#endif
struct alignas(16) FpS {
    uint32_t l[COOP_FPS_WORDS];
};
static_assert(COOP_FPS_WORDS >= 16 && COOP_FPS_WORDS % 4 == 0, "a wide value is four 128-bit words");
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
struct alignas(8) CoopCode8 {
    uint32_t x, y;  // eight int8 term codes, zero-padded
};
static_assert(COOP_SQR_NPROD == COOP_LINE_NPROD, "coop_run2 takes one product count for squares and line products");
struct alignas(16) CoopTables {
    int16_t off[4][3][56];   // [op][x/y/o][...]
    int8_t xt[4][160], yt[4][160], ot[4][336];
    uint32_t c2x[32], c2y[32], c2o[12][2];  // second form of the cyclotomic square (coop_cyc2): packed term codes
    int8_t w2[32];                           // and the weight of every product
    CoopCode8 px[4][56], py[4][56], po[4][60];   // coop_run2: eight packed term codes per input sum / partial output sum
};

// every line of both pairs evaluated at the (scaled) points: A.c0*s, A.c1*s, B.c0*xs, B.c1*xs, ys.  One store per
// machine in the one- and two-machine kernels; ONE store shared by the four machines of pairing_check_kernel.
struct CoopLines {
    FpS lv[2][MILLER_LINES][5];
};
struct CoopWS {
    CoopTables tb;
    FpS prod[54], nprod[54];  // the products of the running operation and their negatives (p - v)
    FpS part[12][5];          // partial output sums (5 lanes per output coefficient)
    FpS reg[COOP_NREG][12], nreg[COOP_NREG][12];  // registers and their negatives (64 p - v)
    CoopLines* ln;   // set by the kernel (coop_attach_lines) before coop_prepare_all_lines
    FpW pt[2][3];    // per pair: s = ZZ*ZZZ and xs = X*ZZZ (times 2^512: see coop_load_points), ys = Y*ZZ (wide)
    FpS onew;        // the constant one, second operand of the cyclotomic square's pass-through products
    FpS zerow;       // padding operand of the sum loops
    Fp canon[12];    // scratch for the excursions into the 12-limb tower code
    Fp2 scr2[8];     // scratch of the cooperative Fp12 inversion (Fp6 cofactors, their products, the Fp2 inverse)
    int use[2];
    int result;
    int cyc2;        // 1: coop_cyc runs the second form of the cyclotomic square (default; 0 = the general schedule, A/B)
    int run2;        // 1: products / squares / line products run coop_run2 (packed codes, tree sums; default), 0: coop_run
    long long* ticks;  // measurement hook (ckzg_b200_debug_pairing_probe): clock64() at the marks below; nullptr in every product kernel
};
#if KZG_DEVICE_PATH
#define COOP_TICK(ws, i)                                                               \
    do {                                                                               \
        if ((ws).ticks && (threadIdx.x & (COOP_LANES - 1)) == 0) (ws).ticks[i] = clock64(); \
    } while (0)
#else
#define COOP_TICK(ws, i) \
    do {                 \
    } while (0)
#endif
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
// packed forms for coop_run2: the (up to eight, zero-padded) term codes of every input sum, and of the partial
// output sum of lane 5 k + sub (terms sub, sub + 5, ... of output coefficient k)
KZG_HD CoopCode8 coop_pack8(const int8_t* terms, int first, int end, int stride) {
    uint32_t c[2] = {0, 0};
    int k = 0;
    for (int t = first; t < end && k < 8; t += stride, k++) c[k >> 2] |= (uint32_t)(uint8_t)terms[t] << (8 * (k & 3));
    CoopCode8 r;
    r.x = c[0];
    r.y = c[1];
    return r;
}
KZG_HD void coop_pack_table(CoopTables& tb, int op, int lane, int nprod, const int16_t* xo, const int16_t* yo, const int16_t* oo, const int8_t* xt, const int8_t* yt,
                            const int8_t* ot) {
    for (int i = lane; i < nprod; i += COOP_LANES) {
        tb.px[op][i] = coop_pack8(xt, xo[i], xo[i + 1], 1);
        tb.py[op][i] = coop_pack8(yt, yo[i], yo[i + 1], 1);
    }
    for (int i = lane; i < 60; i += COOP_LANES) tb.po[op][i] = coop_pack8(ot, oo[i / 5] + i % 5, oo[i / 5 + 1], 5);
}
// which line store this machine reads (and fills its share of); once, after coop_init_tables
KZG_HD void coop_attach_lines(CoopWS& ws, CoopLines* ln) {
    COOP_BEGIN
    if (lane == 0) ws.ln = ln;
    COOP_END
}
// must run once before any coop_run
KZG_HD void coop_init_tables(CoopWS& ws) {
    COOP_BEGIN
    coop_pack_table(ws.tb, COOP_OP_MUL, lane, COOP_MUL_NPROD, COOP_MUL_XOFF, COOP_MUL_YOFF, COOP_MUL_OOFF, COOP_MUL_XT, COOP_MUL_YT, COOP_MUL_OT);
    coop_pack_table(ws.tb, COOP_OP_SQR, lane, COOP_SQR_NPROD, COOP_SQR_XOFF, COOP_SQR_YOFF, COOP_SQR_OOFF, COOP_SQR_XT, COOP_SQR_YT, COOP_SQR_OT);
    coop_pack_table(ws.tb, COOP_OP_LINE, lane, COOP_LINE_NPROD, COOP_LINE_XOFF, COOP_LINE_YOFF, COOP_LINE_OOFF, COOP_LINE_XT, COOP_LINE_YT, COOP_LINE_OT);
    if (lane == 4) ws.run2 = 1;
    coop_copy_table(ws.tb, COOP_OP_MUL, lane, COOP_MUL_NPROD, COOP_MUL_XOFF, COOP_MUL_YOFF, COOP_MUL_OOFF, COOP_MUL_XT, COOP_MUL_YT, COOP_MUL_OT);
    coop_copy_table(ws.tb, COOP_OP_SQR, lane, COOP_SQR_NPROD, COOP_SQR_XOFF, COOP_SQR_YOFF, COOP_SQR_OOFF, COOP_SQR_XT, COOP_SQR_YT, COOP_SQR_OT);
    coop_copy_table(ws.tb, COOP_OP_LINE, lane, COOP_LINE_NPROD, COOP_LINE_XOFF, COOP_LINE_YOFF, COOP_LINE_OOFF, COOP_LINE_XT, COOP_LINE_YT, COOP_LINE_OT);
    coop_copy_table(ws.tb, COOP_OP_CYC, lane, COOP_CYC_NPROD, COOP_CYC_XOFF, COOP_CYC_YOFF, COOP_CYC_OOFF, COOP_CYC_XT, COOP_CYC_YT, COOP_CYC_OT);
    if (lane == 0) s_store(&ws.onew, w_one());
    if (lane == 1) s_store(&ws.zerow, FpW::zero());
    if (lane == 2) ws.ticks = nullptr;
    if (lane == 3) ws.cyc2 = 1;
    if (lane < COOP_CYC_NPROD) {
        uint32_t cx = 0, cy = 0;
        for (int t = COOP_CYC_XOFF[lane]; t < COOP_CYC_XOFF[lane + 1]; t++) cx |= (uint32_t)(uint8_t)COOP_CYC_XT[t] << (8 * (t - COOP_CYC_XOFF[lane]));
        for (int t = COOP_CYC_YOFF[lane]; t < COOP_CYC_YOFF[lane + 1]; t++) cy |= (uint32_t)(uint8_t)COOP_CYC_YT[t] << (8 * (t - COOP_CYC_YOFF[lane]));
        ws.tb.c2x[lane] = cx;
        ws.tb.c2y[lane] = cy;
        ws.tb.w2[lane] = COOP_CYC2_W[lane];
    }
    if (lane < 12) {
        uint32_t c[2] = {0, 0};
        for (int t = COOP_CYC2_OOFF[lane]; t < COOP_CYC2_OOFF[lane + 1]; t++) {
            const int k = t - COOP_CYC2_OOFF[lane];
            c[k >> 2] |= (uint32_t)(uint8_t)COOP_CYC2_OT[t] << (8 * (k & 3));
        }
        ws.tb.c2o[lane][0] = c[0];
        ws.tb.c2o[lane][1] = c[1];
    }
    COOP_END
}
#if KZG_DEVICE_PATH
// The same state from a copy of the tables built once at setup (pairing_tables_kernel): building them costs 50 us of
// divergent constant-memory reads per kernel (profiles/pairing_probe_R3c.log "inputs"), copying 10 KB does not.
__device__ __forceinline__ void coop_init_from(CoopWS& ws, const CoopTables* built) {
    COOP_BEGIN
    static_assert(sizeof(CoopTables) % 16 == 0, "copied as 128-bit words");
    const uint4* src = reinterpret_cast<const uint4*>(built);
    uint4* dst = reinterpret_cast<uint4*>(&ws.tb);
    for (int i = lane; i < (int)(sizeof(CoopTables) / 16); i += COOP_LANES) dst[i] = src[i];
    if (lane == 0) s_store(&ws.onew, w_one());
    if (lane == 1) s_store(&ws.zerow, FpW::zero());
    if (lane == 2) ws.ticks = nullptr;
    if (lane == 3) ws.cyc2 = 1;
    if (lane == 4) ws.run2 = 1;
    COOP_END
}
#endif
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
        const FpW pr = w_mul(x, y);  // < p (+ 2^376)
        s_store(&ws.prod[L], pr);
        FpW npr;
        limbs_sub<14>(npr.l, FPW_CONST_PNEG, pr.l);  // p - v in (0, p], or 2 p - v
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

// ------------------------------------------------------------------------------------------------
// The cyclotomic square, second form.  315 of the ~490 dependent operations of a pairing check are cyclotomic squares
// (five exponentiations by the curve parameter), 3.3 us each in the general schedule above (profiles/
// pairing_probe_R3a.log); what they cost beyond the one Montgomery product is the sums.  Here:
//  * every product of the Granger-Scott formulas enters the outputs with ONE weight (3, 6, or 2 for the pass-through
//    products z * 1), so the lane that owns the product scales it (a 14-step multiply-add chain) and the output forms
//    become plain sums of at most 7 stored values -- one pass of 12 lanes instead of 60 partial sums, a barrier and a
//    second gathering pass (213 unit terms -> 54);
//  * all sums are trees of independent additions started from loaded values (no accumulation into zero): the carry
//    chains of different terms overlap in the pipeline.
// Bounds (tools/gen_pairing_tables.py): scaled products < 6 p, stored negatives 8 p - v, outputs <= 43 p < 64 p.
// ------------------------------------------------------------------------------------------------
KZG_HD FpW coop_load_term(int code, const FpS* a, const FpS* na, const FpS* b, const FpS* zero) {
    const int idx = (code < 0 ? -code : code) - 1;
    const FpS* src = (code == 0) ? zero : (idx < 12 ? ((code < 0 ? na : a) + idx) : (b + (idx - 12)));
    return s_load(src);
}
KZG_HD FpW w_small_mul(const FpW& v, uint32_t w) {  // v * w, w < 8, no overflow for v < 2^445
    FpW r;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
        const uint64_t t = (uint64_t)v.l[i] * w + c;
        r.l[i] = (uint32_t)t;
        c = t >> 32;
    }
    return r;
}
// term codes travel four to a 32-bit word (one shared-memory load per sum, decoded from registers: a load of the code
// followed by a dependent load of the value, once per term, was most of what a sum cost)
KZG_HD int coop_code_at(uint32_t packed, int t) { return (int)(int8_t)(packed >> (8 * t)); }
KZG_HD_NOINLINE void coop_cyc2(CoopWS& ws, int d, int ra) {
    const FpS *a = ws.reg[ra], *na = ws.nreg[ra], *b = &ws.onew, *zero = &ws.zerow;
    COOP_BEGIN
    if (lane < COOP_CYC_NPROD) {
        const uint32_t cx = ws.tb.c2x[lane], cy = ws.tb.c2y[lane];  // 1, 2 or 4 terms each, zero-padded
        FpW x = coop_load_term(coop_code_at(cx, 0), a, na, b, zero), y = coop_load_term(coop_code_at(cy, 0), a, na, b, zero);
        FpW x1 = coop_load_term(coop_code_at(cx, 1), a, na, b, zero), y1 = coop_load_term(coop_code_at(cy, 1), a, na, b, zero);
        FpW x2 = coop_load_term(coop_code_at(cx, 2), a, na, b, zero), y2 = coop_load_term(coop_code_at(cy, 2), a, na, b, zero);
        FpW x3 = coop_load_term(coop_code_at(cx, 3), a, na, b, zero), y3 = coop_load_term(coop_code_at(cy, 3), a, na, b, zero);
        w_acc(x, x1);
        w_acc(y, y1);
        w_acc(x2, x3);
        w_acc(y2, y3);
        w_acc(x, x2);
        w_acc(y, y2);
        const FpW pr = w_small_mul(w_mul(x, y), (uint32_t)ws.tb.w2[lane]);  // < 6.3 p
        s_store(&ws.prod[lane], pr);
        FpW npr;
        limbs_sub<14>(npr.l, FPW_OFF8, pr.l);  // 8 p - v in (2 p, 8 p]
        s_store(&ws.nprod[lane], npr);
    }
    COOP_END
    COOP_BEGIN
    if (lane < 12) {
        const uint32_t c0 = ws.tb.c2o[lane][0], c1 = ws.tb.c2o[lane][1];  // 4 or 7 terms, zero-padded to 8
        const FpS *prod = ws.prod, *nprod = ws.nprod;
        auto term = [&](uint32_t packed, int t) -> FpW {
            const int code = coop_code_at(packed, t);
            const int idx = (code < 0 ? -code : code) - 1;
            return s_load(code == 0 ? zero : ((code < 0 ? nprod : prod) + idx));
        };
        FpW u0 = term(c0, 0), u1 = term(c0, 1), u2 = term(c0, 2), u3 = term(c0, 3);
        w_acc(u0, term(c1, 0));
        w_acc(u1, term(c1, 1));
        w_acc(u2, term(c1, 2));
        w_acc(u0, u1);
        w_acc(u2, u3);
        w_acc(u0, u2);
        coop_put(ws, d, lane, u0);  // <= 43 p, and 64 p - u
    }
    COOP_END
}

// coop_run with the term codes packed eight to a pair of words (one load per sum, decoded from registers) and every
// sum a tree of independent additions: the general schedule above pays a code load, a dependent value load and a
// 14-limb carry chain per term, one term after the other.  Same products, same values, same bounds.
KZG_HD FpW coop_sum8_in(CoopCode8 c, const FpS* a, const FpS* na, const FpS* b, const FpS* zero) {
    FpW s0 = coop_load_term(coop_code_at(c.x, 0), a, na, b, zero), s1 = coop_load_term(coop_code_at(c.x, 1), a, na, b, zero);
    FpW s2 = coop_load_term(coop_code_at(c.x, 2), a, na, b, zero), s3 = coop_load_term(coop_code_at(c.x, 3), a, na, b, zero);
    w_acc(s0, coop_load_term(coop_code_at(c.y, 0), a, na, b, zero));
    w_acc(s1, coop_load_term(coop_code_at(c.y, 1), a, na, b, zero));
    w_acc(s2, coop_load_term(coop_code_at(c.y, 2), a, na, b, zero));
    w_acc(s3, coop_load_term(coop_code_at(c.y, 3), a, na, b, zero));
    w_acc(s0, s1);
    w_acc(s2, s3);
    w_acc(s0, s2);
    return s0;
}
KZG_HD FpW coop_load_prod(int code, const FpS* prod, const FpS* nprod, const FpS* zero) {
    const int idx = (code < 0 ? -code : code) - 1;
    return s_load(code == 0 ? zero : ((code < 0 ? nprod : prod) + idx));
}
KZG_HD_NOINLINE void coop_run2(CoopWS& ws, int op, int d, int ra, const FpS* b) {
    const FpS *a = ws.reg[ra], *na = ws.nreg[ra], *zero = &ws.zerow;
    const int nprod = op == COOP_OP_MUL ? COOP_MUL_NPROD : COOP_SQR_NPROD;  // sqr and line: 36 both
    COOP_BEGIN
    if (lane < nprod) {
        const FpW x = coop_sum8_in(ws.tb.px[op][lane], a, na, b, zero);
        const FpW y = coop_sum8_in(ws.tb.py[op][lane], a, na, b, zero);
        const FpW pr = w_mul(x, y);  // < p (+ 2^376)
        s_store(&ws.prod[lane], pr);
        FpW npr;
        limbs_sub<14>(npr.l, FPW_CONST_PNEG, pr.l);  // p - v in (0, p], or 2 p - v
        s_store(&ws.nprod[lane], npr);
    }
    COOP_END
    COOP_BEGIN
    if (lane < 60) {
        const CoopCode8 c = ws.tb.po[op][lane];
        const FpS *prod = ws.prod, *nprod_s = ws.nprod;
        FpW s0 = coop_load_prod(coop_code_at(c.x, 0), prod, nprod_s, zero), s1 = coop_load_prod(coop_code_at(c.x, 1), prod, nprod_s, zero);
        FpW s2 = coop_load_prod(coop_code_at(c.x, 2), prod, nprod_s, zero), s3 = coop_load_prod(coop_code_at(c.x, 3), prod, nprod_s, zero);
        w_acc(s0, coop_load_prod(coop_code_at(c.y, 0), prod, nprod_s, zero));
        w_acc(s1, coop_load_prod(coop_code_at(c.y, 1), prod, nprod_s, zero));
        w_acc(s2, coop_load_prod(coop_code_at(c.y, 2), prod, nprod_s, zero));
        w_acc(s3, coop_load_prod(coop_code_at(c.y, 3), prod, nprod_s, zero));
        w_acc(s0, s1);
        w_acc(s2, s3);
        w_acc(s0, s2);
        s_store(&ws.part[lane / 5][lane % 5], s0);  // <= 8 p
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
KZG_HD void coop_mul(CoopWS& ws, int d, int a, int b) {
    if (ws.run2) coop_run2(ws, COOP_OP_MUL, d, a, ws.reg[b]);
    else coop_run(ws, COOP_OP_MUL, d, a, ws.reg[b]);
}
KZG_HD void coop_sqr(CoopWS& ws, int d, int a) {
    if (ws.run2) coop_run2(ws, COOP_OP_SQR, d, a, ws.reg[a]);
    else coop_run(ws, COOP_OP_SQR, d, a, ws.reg[a]);
}
KZG_HD void coop_cyc(CoopWS& ws, int d, int a) {
    if (ws.cyc2) coop_cyc2(ws, d, a);
    else coop_run(ws, COOP_OP_CYC, d, a, &ws.onew);
}
KZG_HD void coop_line(CoopWS& ws, int d, int a, int pair, int k) {
    if (ws.run2) coop_run2(ws, COOP_OP_LINE, d, a, ws.ln->lv[pair][k]);
    else coop_run(ws, COOP_OP_LINE, d, a, ws.ln->lv[pair][k]);
}

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
    if (lane < 12) coop_put(ws, d, lane, (lane == 0) ? w_one() : FpW::zero());
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

// The one inversion of the final exponentiation, cooperatively (the serial version above spends 200 us on lane 0,
// 10 % of a pairing check: ~110 dependent products of the 12-limb tower code around one Fp inversion):
//   conj(a) -> reg[y];  N = a * conj(a) in Fp6 (one cooperative product; odd powers of w vanish);
//   N^-1 by the cofactor formula of f6_inv with the three cofactors on three lanes;  a^-1 = conj(a) * N^-1.
// What stays serial is the Fp inversion itself (binary Euclid, ~75 us) and ~12 dependent products.
// d, y, t: distinct registers, all different from a.
KZG_HD_NOINLINE void coop_inv_norm(CoopWS& ws, int d, int a, int y, int t) {
    coop_conj(ws, y, a);
    coop_mul(ws, t, a, y);
    COOP_BEGIN
    if (lane < 6) {  // n0, n1, n2 = coefficients of w^0, w^2, w^4
        const int idx = 4 * (lane >> 1) + (lane & 1);
        ws.canon[lane] = w_to_fp(s_load(&ws.reg[t][idx]));
    }
    COOP_END
    COOP_BEGIN
    if (lane < 3) {
        Fp2 n0, n1, n2;
        n0.c0 = ws.canon[0]; n0.c1 = ws.canon[1];
        n1.c0 = ws.canon[2]; n1.c1 = ws.canon[3];
        n2.c0 = ws.canon[4]; n2.c1 = ws.canon[5];
        Fp2 c;
        if (lane == 0) c = f2_sub(f2_sqr(n0), f2_mul_xi(f2_mul(n1, n2)));
        else if (lane == 1) c = f2_sub(f2_mul_xi(f2_sqr(n2)), f2_mul(n0, n1));
        else c = f2_sub(f2_sqr(n1), f2_mul(n0, n2));
        ws.scr2[lane] = c;
        ws.scr2[3 + lane] = f2_mul(lane == 0 ? n0 : (lane == 1 ? n2 : n1), c);  // n0 c0, n2 c1, n1 c2
    }
    COOP_END
    COOP_BEGIN
    if (lane == 0) ws.scr2[6] = f2_inv(f2_add(ws.scr2[3], f2_mul_xi(f2_add(ws.scr2[4], ws.scr2[5]))));
    COOP_END
    COOP_BEGIN
    if (lane < 3) ws.scr2[lane] = f2_mul(ws.scr2[lane], ws.scr2[6]);
    COOP_END
    COOP_BEGIN
    if (lane < 12) {
        const int k = lane >> 1;
        FpW v = FpW::zero();
        if ((k & 1) == 0) v = w_from_fp((lane & 1) ? ws.scr2[k >> 1].c1 : ws.scr2[k >> 1].c0);
        coop_put(ws, d, lane, v);
    }
    COOP_END
    coop_mul(ws, d, y, d);
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
        ws.pt[pair][j] = (j == 2) ? w_from_fp(v) : w_mul(w_ext(v), FpW::from_limbs(FPW_CONST_CPT));
    }
    if (lane == 6) ws.use[0] = (!g1_is_inf(P1) && !L1->is_inf) ? 1 : 0;
    if (lane == 7) ws.use[1] = (!g1_is_inf(P2) && !L2->is_inf) ? 1 : 0;
    COOP_END
}

// Every line of both pairs evaluated at the (scaled) points -> ws.lv, in one parallel pass before the
// Miller loop (544 independent products over the 64 lanes).  The first version did this step by step
// inside the loop: 68 extra barriers and 68 exposed global-memory round trips, 15 % of the kernel.
// `part` of `nparts`: machines that share one line store split the work (each still ends at its own barrier; the
// caller synchronises the machines before any of them reads the store)
KZG_HD_NOINLINE void coop_prepare_all_lines(CoopWS& ws, const G2Lines* L1, const G2Lines* L2, int part = 0, int nparts = 1) {
    COOP_BEGIN
    for (int i = part * COOP_LANES + lane; i < 2 * MILLER_LINES * 5; i += nparts * COOP_LANES) {
        const int pair = i / (MILLER_LINES * 5), r = i % (MILLER_LINES * 5), k = r / 5, j = r % 5;
        if (!ws.use[pair]) continue;
        const LineCoeff& l = (pair ? L2 : L1)->line[k];
        FpW v;
        if (j == 0) v = w_mul(w_ext(l.A.c0), ws.pt[pair][0]);
        else if (j == 1) v = w_mul(w_ext(l.A.c1), ws.pt[pair][0]);
        else if (j == 2) v = w_mul(w_ext(l.B.c0), ws.pt[pair][1]);
        else if (j == 3) v = w_mul(w_ext(l.B.c1), ws.pt[pair][1]);
        else v = ws.pt[pair][2];
        s_store(&ws.ln->lv[pair][k][j], v);
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
KZG_HD void coop_pairing_product_is_one(CoopWS& ws, CoopLines* ln, const G1& P1, const G2Lines* L1, const G1& P2, const G2Lines* L2, bool negate_first,
                                        const CoopTables* built = nullptr) {
#if KZG_DEVICE_PATH
    if (built) coop_init_from(ws, built);
    else
#endif
        coop_init_tables(ws);
    (void)built;
    coop_attach_lines(ws, ln);
    coop_load_points(ws, P1, L1, P2, L2, negate_first);
    coop_prepare_all_lines(ws, L1, L2);
    coop_miller_loop(ws);
    coop_final_exp_is_one(ws);
}

// ws.result = [ reg[F]^((p^12 - 1) / r) == 1 ]
KZG_HD void coop_final_exp_is_one(CoopWS& ws) {
    enum { F = 0, E = 1, T0 = 2, T1 = 3, T2 = 4, T3 = 5, X = 6, Y = 7 };
    // ---- final exponentiation, easy part: E = F^((p^6-1)(p^2+1)) ----
    COOP_TICK(ws, 4);
    coop_inv(ws, X, F);
    COOP_TICK(ws, 5);
    coop_conj(ws, Y, F);
    coop_mul(ws, E, Y, X);
    coop_frobenius(ws, X, E, 2);
    coop_mul(ws, E, X, E);
    COOP_TICK(ws, 6);
    // ---- hard part: E^((z-1)^2 (z+p)(z^2+p^2-1) + 3) ----
    coop_pow_x(ws, T0, E);            // E^z
    COOP_TICK(ws, 7);
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
    COOP_TICK(ws, 8);
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

// ------------------------------------------------------------------------------------------------
// The final exponentiation on TWO machines (pairing_check_kernel keeps machine 1 alive after its Miller loop).
// An exponentiation by the curve parameter is 63 dependent cyclotomic squarings; the 5 products of square-and-multiply
// and the factor that follows every exponentiation in the hard part (a conjugate, a Frobenius image, E^3) need not be
// on that chain: machine 0 only squares, right to left, and PUBLISHES s^(2^b) into machine 1's registers at the set
// bits of |z| (16, 48, 57, 60, 62, 63); machine 1 starts its accumulator at conj(factor) and multiplies the published
// powers in as they arrive:  conj(conj(factor) * s^|z|) = factor * s^z.  Per exponentiation the chain is 63 squarings
// plus one product (the last published power) instead of 63 squarings, 6 products, a conjugation and a Frobenius.
// Both machines meet at CTA-wide named barrier 3; on the host (tests/hostcheck) the roles run one after the other.
// ------------------------------------------------------------------------------------------------
// reg[d] of `dst` := reg[a] of `src` (another machine's workspace), optionally conjugated
KZG_HD_NOINLINE void coop_copy_x(CoopWS& dst, int d, const CoopWS& src, int a, bool conj) {
#if KZG_DEVICE_PATH
    {
        const int lane = threadIdx.x & (COOP_LANES - 1);
#else
    for (int lane = 0; lane < COOP_LANES; lane++) {
#endif
        if (lane < 12) {
            const bool sw = conj && ((lane >> 1) & 1);
            const FpW v = s_load(&src.reg[a][lane]), nv = s_load(&src.nreg[a][lane]);
            s_store(&dst.reg[d][lane], sw ? nv : v);
            s_store(&dst.nreg[d][lane], sw ? v : nv);
        }
    }
}
// Named barriers: 1 + m = machine m's own (64 threads), 8 = machines 0 and 1 (128 threads), 8 + h = machine 0 and
// helper h of the four-machine Miller loop (128 threads).
#if KZG_DEVICE_PATH
__device__ __forceinline__ void duo_sync() { asm volatile("bar.sync 8, 128;" ::: "memory"); }
__device__ __forceinline__ void quad_sync(int h) { asm volatile("bar.sync %0, 128;" ::"r"(8 + h) : "memory"); }
#define DUO_IS(r) (g == (r))
#else
inline void duo_sync() {}
inline void quad_sync(int) {}
#define DUO_IS(r) true
#endif

// ------------------------------------------------------------------------------------------------
// The Miller loop on FOUR machines.  f <- f^2 * l_i (* l'_i) is a chain of 63 squarings and 136 line products in
// the textbook loop; two machines (one pairing each) made it 63 + 68.  But the lines do not depend on f:
//     f_after = f_before^(2^g) * M,     M = the same loop over g iterations started from 1,
// so three HELPER machines run the loop over groups of iterations (both pairings' lines, shared squarings) and the
// CHAIN machine only squares and multiplies one full Fp12 value per group in: 61 squarings + 13 products instead of
// 131 operations (groups of 2, 3, 4, 4, then 5 iterations: small at first so that the chain starts after 26 us; a
// helper needs ~50 us per group of five and the chain asks each for one every 59 us).  Helper h alternates between two
// registers of its own for its results; the chain multiplies straight out of the helper's shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int MQ_GROUPS = 14;
KZG_CONST int8_t MQ_START[MQ_GROUPS + 1] = {0, 2, 5, 9, 13, 18, 23, 28, 33, 38, 43, 48, 53, 58, 63};
// w[0] = chain, w[1..3] = helpers; g = the calling thread's machine.  Result: w[0].reg[0] = conj(Miller value of both pairs).
KZG_HD void coop_miller_quad(CoopWS* w, int g) {
    (void)g;
    enum { F = 0 };
    const uint64_t z = BLS_X_ABS;
    for (int j = 0; j < MQ_GROUPS; j++) {
        const int h = 1 + j % 3, M = 2 + ((j / 3) & 1);
        const int i0 = MQ_START[j], i1 = MQ_START[j + 1];
        if (DUO_IS(h)) {
            CoopWS& ws = w[h];
            const bool use0 = ws.use[0] != 0, use1 = ws.use[1] != 0;
            int k = i0;  // index of iteration i0's first line: one per iteration plus one per set bit above it
            for (int b = 62; b > 62 - i0; b--) k += (int)((z >> b) & 1ull);
            coop_set_one(ws, M);
            for (int i = i0; i < i1; i++) {
                const int b = 62 - i;
                if (i != i0) coop_sqr(ws, M, M);
                if (use0) coop_line(ws, M, M, 0, k);
                if (use1) coop_line(ws, M, M, 1, k);
                k++;
                if ((z >> b) & 1ull) {
                    if (use0) coop_line(ws, M, M, 0, k);
                    if (use1) coop_line(ws, M, M, 1, k);
                    k++;
                }
            }
        }
        if (DUO_IS(h) || DUO_IS(0)) quad_sync(h);
        if (DUO_IS(0)) {
            CoopWS& ws = w[0];
            if (j == 0) {
                coop_copy_x(ws, F, w[h], M, false);
                coop_sync();
            } else {
                for (int i = i0; i < i1; i++) coop_sqr(ws, F, F);
                if (ws.run2) coop_run2(ws, COOP_OP_MUL, F, F, w[h].reg[M]);
                else coop_run(ws, COOP_OP_MUL, F, F, w[h].reg[M]);
            }
        }
    }
    if (DUO_IS(0)) coop_conj(w[0], F, F);
}
enum { DUO_ACC = 0, DUO_SLOT0 = 1, DUO_NSLOT = 6 };  // machine 1: accumulator, published powers (registers 1..6), 7 = scratch

// machine 1: accumulator := conj(factor of this exponentiation); reads machine 0's registers (complete: duo_sync before)
KZG_HD void duo_acc_init(CoopWS& w1, const CoopWS& w0, int kind, int s, int rE, int rT2) {
    if (kind == 0) {  // factor conj(s): the (z - 1) steps
        coop_copy_x(w1, DUO_ACC, w0, s, false);
        coop_sync();
    } else if (kind == 1) {  // factor frob(s): the (z + p) step
        coop_copy_x(w1, 7, w0, s, false);
        coop_sync();
        coop_frobenius(w1, 6, 7, 1);
        coop_conj(w1, DUO_ACC, 6);
    } else if (kind == 2) {  // no factor
        coop_set_one(w1, DUO_ACC);
    } else {  // factor frob2(T2) * conj(T2) * E^3: the (z^2 + p^2 - 1) step and the final + 3
        coop_copy_x(w1, 7, w0, rE, false);
        coop_sync();
        coop_cyc(w1, 5, 7);
        coop_mul(w1, 5, 5, 7);  // E^3
        coop_copy_x(w1, 7, w0, rT2, false);
        coop_sync();
        coop_frobenius(w1, 6, 7, 2);
        coop_conj(w1, 7, 7);
        coop_mul(w1, 6, 6, 7);
        coop_mul(w1, 6, 6, 5);
        coop_conj(w1, DUO_ACC, 6);
    }
}
// reg[d] of machine 0 = factor(kind) * reg[s]^z;  d != s
KZG_HD_NOINLINE void coop_pow_x_duo(CoopWS& w0, CoopWS& w1, int g, int d, int s, int kind, int rE, int rT2) {
    (void)g;
    const uint64_t z = BLS_X_ABS;
    duo_sync();  // machine 0's registers are complete
    if (DUO_IS(1)) duo_acc_init(w1, w0, kind, s, rE, rT2);
    if (DUO_IS(0)) {
        coop_copy(w0, d, s);
        int slot = 0;
        for (int b = 0; b < 64; b++) {
            if ((z >> b) & 1ull) {
                coop_copy_x(w1, DUO_SLOT0 + slot, w0, d, false);
                duo_sync();
                slot++;
            }
            if (b < 63) coop_cyc(w0, d, d);
        }
    }
    if (DUO_IS(1))
        for (int slot = 0; slot < DUO_NSLOT; slot++) {
            duo_sync();
            coop_mul(w1, DUO_ACC, DUO_ACC, DUO_SLOT0 + slot);
        }
    duo_sync();  // the accumulator is complete
    if (DUO_IS(0)) {
        coop_copy_x(w0, d, w1, DUO_ACC, true);
        coop_sync();
    }
}
// w0.result = [ w0.reg[0]^((p^12 - 1) / r) == 1 ];  every thread of both machines calls this (g = its machine)
KZG_HD void coop_final_exp_is_one_duo(CoopWS& w0, CoopWS& w1, int g) {
    enum { F = 0, E = 1, T0 = 2, T1 = 3, T2 = 4, T3 = 5, X = 6, Y = 7 };
    if (DUO_IS(0)) {
        COOP_TICK(w0, 4);
        coop_inv_norm(w0, X, F, Y, T0);   // X = 1 / F, Y = conj(F)
        COOP_TICK(w0, 5);
        coop_mul(w0, E, Y, X);
        coop_frobenius(w0, X, E, 2);
        coop_mul(w0, E, X, E);            // E = F^((p^6 - 1)(p^2 + 1))
        COOP_TICK(w0, 6);
    }
    // hard part: E^((z-1)^2 (z+p)(z^2+p^2-1) + 3)
    coop_pow_x_duo(w0, w1, g, T0, E, 0, E, T2);    // E^(z-1)
    if (DUO_IS(0)) COOP_TICK(w0, 7);
    coop_pow_x_duo(w0, w1, g, T1, T0, 0, E, T2);   // ^(z-1)^2
    coop_pow_x_duo(w0, w1, g, T2, T1, 1, E, T2);   // ^(z+p)
    coop_pow_x_duo(w0, w1, g, X, T2, 2, E, T2);    // T2^z
    coop_pow_x_duo(w0, w1, g, T3, X, 3, E, T2);    // T2^(z^2 + p^2 - 1) * E^3
    if (DUO_IS(0)) {
        COOP_TICK(w0, 8);
        COOP_BEGIN
        if (lane < 12) w0.canon[lane] = w_to_fp(s_load(&w0.reg[T3][lane]));
        COOP_END
        COOP_BEGIN
        if (lane == 0) {
            bool one = eq(w0.canon[0], Fp::one());
            for (int i = 1; i < 12; i++) one = one && is_zero(w0.canon[i]);
            w0.result = one ? 1 : 0;
        }
        COOP_END
    }
}

}  // namespace kzg
