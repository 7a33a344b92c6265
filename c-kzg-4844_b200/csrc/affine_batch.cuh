// One pair of a batched affine addition (msm_affine.cu): what it contributes to the shared inversion and how its
// sum is finished once 1 / d is known.  Host-compilable (tests/hostcheck pins every branch -- generic sum,
// doubling, opposite points, infinity on either side -- against the integer oracle).
#pragma once
#include "g1.cuh"

namespace kzg {

// What one pair contributes to the batched inversion, and how its sum is finished afterwards.
enum { BAM_ADD = 0, BAM_DBL = 1, BAM_TAKE_A = 2, BAM_TAKE_B = 3, BAM_INF = 4 };

// denominator of the pair (never zero) and the kind of the pair.  ya / yb are loaded only when needed.
template <class LoadYA, class LoadYB>
KZG_HD int bam_plan(Fp& d, const Fp& xa, const Fp& xb, bool inf_a, bool inf_b, LoadYA load_ya, LoadYB load_yb) {
    d = Fp::one();
    if (inf_a && inf_b) return BAM_INF;
    if (inf_b) return BAM_TAKE_A;
    if (inf_a) return BAM_TAKE_B;
    const Fp t = sub(xb, xa);
    if (!is_zero(t)) {
        d = t;
        return BAM_ADD;
    }
    const Fp ya = load_ya(), yb = load_yb();
    if (eq(ya, yb) && !is_zero(ya)) {  // the same point: doubling, slope 3 x^2 / (2 y)
        d = dbl(ya);
        return BAM_DBL;
    }
    return BAM_INF;  // opposite points (or a 2-torsion point, which the subgroup does not contain)
}

// the sum of the pair given 1 / d
KZG_HD void bam_finish(Fp& x3, Fp& y3, int kind, const Fp& xa, const Fp& ya, const Fp& xb, const Fp& yb, const Fp& dinv) {
    if (kind == BAM_TAKE_A) {
        x3 = xa;
        y3 = ya;
        return;
    }
    if (kind == BAM_TAKE_B) {
        x3 = xb;
        y3 = yb;
        return;
    }
    if (kind == BAM_INF) {
        x3 = Fp::zero();
        y3 = Fp::zero();
        return;
    }
    Fp num;
    if (kind == BAM_DBL) {
        const Fp xx = sqr(xa);
        num = add(dbl(xx), xx);
    } else {
        num = sub(yb, ya);
    }
    const Fp lam = mul(num, dinv);
    x3 = sub(sub(sqr(lam), xa), xb);
    y3 = sub(mul(lam, sub(xa, x3)), ya);
}

// Reference use of the two halves above, as the kernels string them together: out[i] = a[i] + b[i] for n pairs
// with ONE field inversion (prefix products up, two products per pair down).  n <= 64.
KZG_HD void affine_batch_add(G1Affine* out, const G1Affine* a, const G1Affine* b, int n) {
    Fp prefix[64];
    Fp run = Fp::one();
    for (int i = 0; i < n; i++) {
        Fp d;
        (void)bam_plan(d, a[i].x, b[i].x, g1a_is_inf(a[i]), g1a_is_inf(b[i]), [&] { return a[i].y; }, [&] { return b[i].y; });
        prefix[i] = run;
        run = mul(run, d);
    }
    Fp inv = fp_inv(run);
    for (int i = n - 1; i >= 0; i--) {
        Fp d;
        const int kind = bam_plan(d, a[i].x, b[i].x, g1a_is_inf(a[i]), g1a_is_inf(b[i]), [&] { return a[i].y; }, [&] { return b[i].y; });
        const Fp dinv = mul(inv, prefix[i]);
        inv = mul(inv, d);
        bam_finish(out[i].x, out[i].y, kind, a[i].x, a[i].y, b[i].x, b[i].y, dinv);
    }
}

}  // namespace kzg
