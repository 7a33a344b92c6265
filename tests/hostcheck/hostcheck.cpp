// TEST-ONLY host build of the engine's header-only arithmetic (portable uint64 path of field.cuh).
// Compiled by tests/test_hostcheck.py with g++ into tests/hostcheck/_build/libhostcheck.so and
// compared against the Python oracle.  Never linked into the product library: the product has no
// CPU path.  Purpose: catch formula mistakes in the GPU-less build container before spending GPU time.
#include <string.h>

#include "g1.cuh"
#include "affine_batch.cuh"
#ifdef HOSTCHECK_PAIRING
#include "pairing.cuh"
#include "pairing_coop.cuh"
#endif

using namespace kzg;

extern "C" {

// plain little-endian limbs in / out
void hc_fp_mul(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    Fp x = to_mont<FpTag>(a), y = to_mont<FpTag>(b);
    from_mont<FpTag>(out, mul(x, y));
}
void hc_fp_addsub(uint32_t* out_add, uint32_t* out_sub, uint32_t* out_neg, const uint32_t* a, const uint32_t* b) {
    Fp x = to_mont<FpTag>(a), y = to_mont<FpTag>(b);
    from_mont<FpTag>(out_add, add(x, y));
    from_mont<FpTag>(out_sub, sub(x, y));
    from_mont<FpTag>(out_neg, neg(x));
}
void hc_fp_inv(uint32_t* out, const uint32_t* a) {
    from_mont<FpTag>(out, fp_inv(to_mont<FpTag>(a)));
}
void hc_fr_mul(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    Fr x = to_mont<FrTag>(a), y = to_mont<FrTag>(b);
    from_mont<FrTag>(out, mul(x, y));
}
void hc_fr_inv(uint32_t* out, const uint32_t* a) {
    from_mont<FrTag>(out, fr_inv(to_mont<FrTag>(a)));
}
int hc_fr_from_be(uint8_t* out_be, const uint8_t* in_be) {
    Fr x;
    int ok = fr_from_be_checked(x, in_be);
    fr_to_be(out_be, x);
    return ok;
}
void hc_fr_hash_reduce(uint8_t* out_be, const uint8_t* in_be) {
    fr_to_be(out_be, fr_from_be_reduce(in_be));
}

// 48-byte compressed in -> validate; returns 1 ok / 0 bad; re-compressed point out
int hc_g1_validate(uint8_t* out48, const uint8_t* in48) {
    G1Affine a;
    int ok = g1a_validate(a, in48);
    if (ok) g1a_compress(out48, a);
    return ok;
}
// validation with table levels: 18 compressed points out (levels), returns the verdict
int hc_g1_validate_levels(uint8_t* out48x18, const uint8_t* in48) {
    G1Affine a;
    G1 levels[G1_LEVELS];
    int ok = g1a_validate_levels(a, in48, levels, 1);
    for (int j = 0; j < G1_LEVELS; j++) g1a_compress(out48x18 + 48 * j, g1_to_affine(levels[j]));
    return ok;
}
void hc_basez_split(int64_t* s4, const uint32_t* k8) { basez_split(s4, k8); }
int hc_g1_uncompress(uint8_t* out48, const uint8_t* in48) {
    G1Affine a;
    int ok = g1a_uncompress(a, in48);
    if (ok) g1a_compress(out48, a);
    return ok;
}
// out = [k]P (+ Q if q48 != NULL), everything compressed; k = 8 plain LE limbs
int hc_g1_mul_add(uint8_t* out48, const uint8_t* p48, const uint32_t* k, const uint8_t* q48) {
    G1Affine p, q;
    if (!g1a_uncompress(p, p48)) return 0;
    G1 r = g1_mul_affine<8>(p, k);
    if (q48) {
        if (!g1a_uncompress(q, q48)) return 0;
        G1 qq = g1_from_affine(q);
        r = g1_add(r, qq);
    }
    g1a_compress(out48, g1_to_affine(r));
    return 1;
}
// out = P + Q via madd (P lifted to XYZZ after a doubling-and-back to exercise non-trivial zz)
int hc_g1_madd(uint8_t* out48, const uint8_t* p48, const uint8_t* q48, int negate) {
    G1Affine p, q;
    if (!g1a_uncompress(p, p48) || !g1a_uncompress(q, q48)) return 0;
    G1 acc = g1_from_affine(p);
    // make zz != 1: acc = (P + P) - P
    if (!g1_is_inf(acc)) {
        G1 d = g1_dbl(acc);
        g1_madd(d, p, true);
        acc = d;
    }
    g1_madd(acc, q, negate != 0);
    g1a_compress(out48, g1_to_affine(acc));
    return 1;
}

// out[i] = a[i] + b[i] for n <= 64 pairs of compressed points through the batched affine addition (one inversion)
int hc_affine_batch_add(uint8_t* out48, const uint8_t* a48, const uint8_t* b48, int n) {
    G1Affine a[64], b[64], o[64];
    if (n < 0 || n > 64) return 0;
    for (int i = 0; i < n; i++)
        if (!g1a_uncompress(a[i], a48 + 48 * i) || !g1a_uncompress(b[i], b48 + 48 * i)) return 0;
    affine_batch_add(o, a, b, n);
    for (int i = 0; i < n; i++) g1a_compress(out48 + 48 * i, o[i]);
    return 1;
}

#ifdef HOSTCHECK_PAIRING
// e(P1,Q1) * e(P2,Q2) == 1 ?  (compressed inputs; -1 on decode failure)
int hc_pairing_product_is_one(const uint8_t* p1, const uint8_t* q1, const uint8_t* p2, const uint8_t* q2) {
    G1Affine P1, P2;
    G2Affine Q1, Q2;
    if (!g1a_uncompress(P1, p1) || !g1a_uncompress(P2, p2)) return -1;
    if (!g2a_uncompress(Q1, q1) || !g2a_uncompress(Q2, q2)) return -1;
    static G2Lines L1, L2;
    g2_precompute_lines(L1, Q1);
    g2_precompute_lines(L2, Q2);
    return pairing_product_is_one(P1, &L1, P2, &L2) ? 1 : 0;
}
int hc_g2_uncompress_ok(const uint8_t* q) {
    G2Affine Q;
    return g2a_uncompress(Q, q) ? 1 : 0;
}
// cyclotomic squaring == plain squaring on an element of the cyclotomic subgroup (easy part of a
// Miller-loop value); also x-power consistency.  returns 1 if all agree
int hc_cyclotomic_consistency(const uint8_t* p1, const uint8_t* q1) {
    G1Affine P1;
    G2Affine Q1;
    if (!g1a_uncompress(P1, p1) || !g2a_uncompress(Q1, q1)) return -1;
    static G2Lines L1;
    g2_precompute_lines(L1, Q1);
    G1Affine inf = g1a_inf();
    Fp12 f = miller_loop2(P1, &L1, inf, &L1);
    Fp12 e = f12_mul(f12_conj(f), f12_inv(f));
    e = f12_mul(f12_frobenius2(e), e);
    Fp12 a = f12_cyclotomic_sqr(e), b = f12_sqr(e);
    int ok = f6_eq(a.c0, b.c0) && f6_eq(a.c1, b.c1);
    // conj == inverse in the cyclotomic subgroup
    Fp12 one = f12_mul(e, f12_conj(e));
    ok = ok && f12_is_one(one);
    // frobenius consistency: frob(frob(e)) == frob2(e)
    Fp12 f2a = f12_frobenius(f12_frobenius(e)), f2b = f12_frobenius2(e);
    ok = ok && f6_eq(f2a.c0, f2b.c0) && f6_eq(f2a.c1, f2b.c1);
    return ok;
}
// the thread-cooperative schedule (pairing_coop.cuh), lanes emulated by loops
int hc_coop_pairing_product_is_one(const uint8_t* p1, const uint8_t* q1, const uint8_t* p2, const uint8_t* q2, int negate_first, int scale) {
    G1Affine P1, P2;
    G2Affine Q1, Q2;
    if (!g1a_uncompress(P1, p1) || !g1a_uncompress(P2, p2)) return -1;
    if (!g2a_uncompress(Q1, q1) || !g2a_uncompress(Q2, q2)) return -1;
    static G2Lines L1, L2;
    g2_precompute_lines(L1, Q1);
    g2_precompute_lines(L2, Q2);
    G1 A = g1_from_affine(P1), Bp = g1_from_affine(P2);
    if (scale) {  // exercise non-trivial ZZ/ZZZ: (2P - P) and (P + P - P)
        if (!g1_is_inf(A)) { G1 d = g1_dbl(A); g1_madd(d, P1, true); A = d; }
        if (!g1_is_inf(Bp)) { G1 d = g1_dbl(Bp); g1_madd(d, P2, true); Bp = d; }
    }
    static CoopWS ws;
    static CoopLines ln1;
    coop_pairing_product_is_one(ws, &ln1, A, &L1, Bp, &L2, negate_first != 0);
    const int single = ws.result;
    // the two-machine form of pairing_check_kernel (pairing.cu): one Miller loop per machine, product, final exponentiation
    static CoopWS w2[2];
    static CoopLines ln2;  // one store for both machines: each fills its own pair
    for (int g = 0; g < 2; g++) {
        coop_init_tables(w2[g]);
        coop_attach_lines(w2[g], &ln2);
        w2[g].run2 = 0;  // the general schedule (coop_run) on this path, its packed / tree form (coop_run2) on the others
        w2[g].cyc2 = 0;
        coop_load_points(w2[g], A, &L1, Bp, &L2, negate_first != 0);
        w2[g].use[1 - g] = 0;
        coop_prepare_all_lines(w2[g], &L1, &L2);
        coop_miller_loop(w2[g]);
    }
    for (int lane = 0; lane < 12; lane++) {
        w2[0].reg[6][lane] = w2[1].reg[0][lane];
        w2[0].nreg[6][lane] = w2[1].nreg[0][lane];
    }
    coop_mul(w2[0], 0, 0, 6);
    // the product kernel's final exponentiation: machine 0 squares, machine 1 multiplies (roles run in turn on the host),
    // cooperative inversion; compared with the one-machine form on a copy of the same Miller value
    static CoopWS w3[2];
    w3[0] = w2[0];
    w3[1] = w2[1];
    w3[0].run2 = w3[1].run2 = 1;
    w3[0].cyc2 = w3[1].cyc2 = 1;
    coop_final_exp_is_one(w2[0]);
    if (w2[0].result != single) return -2;
    coop_final_exp_is_one_duo(w3[0], w3[1], 0);
    if (w3[0].result != single) return -3;
    // the four-machine Miller loop of the product kernel (roles in turn on the host) + the two-machine final exponentiation
    static CoopWS w4[4];
    static CoopLines ln4;
    for (int g = 0; g < 4; g++) {
        coop_init_tables(w4[g]);
        coop_attach_lines(w4[g], &ln4);
        coop_load_points(w4[g], A, &L1, Bp, &L2, negate_first != 0);
    }
    for (int g = 0; g < 4; g++) coop_prepare_all_lines(w4[g], &L1, &L2, g, 4);
    coop_miller_quad(w4, 0);
    coop_final_exp_is_one_duo(w4[0], w4[1], 0);
    if (w4[0].result != single) return -5;
    // the cooperative inversion alone: reg[1] * inv(reg[1]) == 1 on the value the easy part left in register 1
    coop_inv_norm(w2[0], 6, 1, 7, 5);
    coop_mul(w2[0], 4, 6, 1);
    for (int i = 0; i < 12; i++) {
        const Fp c = w_to_fp(s_load(&w2[0].reg[4][i]));
        if (i == 0 ? !eq(c, Fp::one()) : !is_zero(c)) return -4;
    }
    return single;
}
#endif
}
