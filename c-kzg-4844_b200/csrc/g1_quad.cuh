// Cooperative XYZZ addition on four lanes (a "quad"), device only.
//
// The tree sums at the end of a verifier's linear combination are pure latency: few additions, each
// waiting for the previous level, and one addition on one lane is 14 dependent Fp products (~14 us).
// add-2008-s has depth four: {U1, U2, S1, S2} -> {PP, RR, ZZ1*ZZ2, ZZZ1*ZZZ2} -> {PPP, Q, ZZ3} ->
// {R*(Q - X3), S1*PPP, ZZZ3}.  Four consecutive lanes take one product of each level; every lane
// picks its operands and then ALL lanes execute the same multiplier call, so the quad pays four
// product latencies instead of fourteen.  Intermediates travel through a 576-byte scratch block in
// shared memory.  Same group law as g1_add (g1.cuh), so the same results.
#pragma once
#include "g1.cuh"

namespace kzg {

struct QuadScratch {
    Fp v[12];
};

__device__ __forceinline__ Fp quad_ld(const Fp* p) {
    Fp a;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = q[i];
    return a;
}
__device__ __forceinline__ void quad_st(Fp* p, const Fp& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* d = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = d[i];
}
__device__ __forceinline__ Fp quad_sel(bool c, const Fp& a, const Fp& b) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}

// complete addition on one lane, for the rare equal-x case (kept out of line: it would otherwise set the
// register budget of the fast path)
static __device__ __noinline__ void g1_add_mem(G1* out, const G1* p, const G1* q) {
    G1 a, b;
    a.x = quad_ld(&p->x); a.y = quad_ld(&p->y); a.zz = quad_ld(&p->zz); a.zzz = quad_ld(&p->zzz);
    b.x = quad_ld(&q->x); b.y = quad_ld(&q->y); b.zz = quad_ld(&q->zz); b.zzz = quad_ld(&q->zzz);
    g1_add_to(a, b);
    quad_st(&out->x, a.x); quad_st(&out->y, a.y); quad_st(&out->zz, a.zz); quad_st(&out->zzz, a.zzz);
}

// *out = *p + *q.  Called by the four lanes of a quad (lanes 4k..4k+3 of a warp) with identical
// arguments; `out` may alias `p` or `q`.  The operands must be visible to all four lanes on entry
// (plain loads: they may have been written earlier in the same kernel); `sc` is private to the quad.
static __device__ __noinline__ void g1_add_quad(G1* out, const G1* p, const G1* q, QuadScratch* sc) {
    const unsigned lane = threadIdx.x & 31u, ql = lane & 3u;
    const unsigned mask = 0xFu << (lane & ~3u);
    const Fp pzz = quad_ld(&p->zz), qzz = quad_ld(&q->zz);
    const bool pinf = is_zero(pzz), qinf = is_zero(qzz);
    if (pinf || qinf) {
        __syncwarp(mask);
        if (ql == 0) {
            const G1* src = pinf ? q : p;
            if (out != src) {
                Fp a = quad_ld(&src->x), b = quad_ld(&src->y), c = quad_ld(&src->zz), d = quad_ld(&src->zzz);
                quad_st(&out->x, a); quad_st(&out->y, b); quad_st(&out->zz, c); quad_st(&out->zzz, d);
            }
        }
        __syncwarp(mask);
        return;
    }
    // level 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
    {
        const Fp* ap = ql == 0 ? &p->x : ql == 1 ? &q->x : ql == 2 ? &p->y : &q->y;
        const Fp* bp = ql == 0 ? &q->zz : ql == 1 ? &p->zz : ql == 2 ? &q->zzz : &p->zzz;
        quad_st(&sc->v[ql], mul(quad_ld(ap), quad_ld(bp)));
    }
    __syncwarp(mask);
    // (U1 and S1 are re-read from the scratch block where needed: fewer live registers)
    const Fp Pd = sub(quad_ld(&sc->v[1]), quad_ld(&sc->v[0])), Rd = sub(quad_ld(&sc->v[3]), quad_ld(&sc->v[2]));
    if (is_zero(Pd)) {  // same x: doubling or cancellation -- rare, one lane does the complete addition
        __syncwarp(mask);
        if (ql == 0) g1_add_mem(out, p, q);  // reads both operands completely before it writes
        __syncwarp(mask);
        return;
    }
    // level 2: PP = P^2, ZZ1 ZZ2, RR = R^2, ZZZ1 ZZZ2
    {
        const Fp* ap = (ql == 1) ? &p->zz : &p->zzz;
        const Fp* bp = (ql == 1) ? &q->zz : &q->zzz;
        const Fp la = quad_ld(ap), lb = quad_ld(bp);
        const Fp d = quad_sel(ql == 0, Pd, Rd);
        const bool even = (ql & 1u) == 0;
        quad_st(&sc->v[4 + ql], mul(quad_sel(even, d, la), quad_sel(even, d, lb)));
    }
    __syncwarp(mask);  // from here on the inputs are dead: `out` may be written even if it aliases them
    // level 3: PPP = P PP, Q = U1 PP, ZZ3 = (ZZ1 ZZ2) PP
    {
        const Fp a = quad_sel(ql == 1, quad_ld(&sc->v[0]), quad_sel(ql == 2, quad_ld(&sc->v[5]), Pd));
        const Fp r3 = mul(a, quad_ld(&sc->v[4]));
        if (ql == 0) quad_st(&sc->v[8], r3);
        if (ql == 1) quad_st(&sc->v[9], r3);
        if (ql == 2) quad_st(&out->zz, r3);
    }
    __syncwarp(mask);
    // level 4: X3 = RR - PPP - 2Q;  R (Q - X3),  S1 PPP,  ZZZ3 = (ZZZ1 ZZZ2) PPP
    const Fp PPP = quad_ld(&sc->v[8]), Qv = quad_ld(&sc->v[9]);
    const Fp X3 = sub(sub(quad_ld(&sc->v[6]), PPP), dbl(Qv));
    Fp r4;
    {
        const Fp a = quad_sel(ql == 1, quad_ld(&sc->v[2]), quad_sel(ql == 2, quad_ld(&sc->v[7]), Rd));
        const Fp b = quad_sel(ql == 0 || ql == 3, sub(Qv, X3), PPP);
        r4 = mul(a, b);
        if (ql == 1) quad_st(&sc->v[10], r4);
        if (ql == 2) quad_st(&out->zzz, r4);
        if (ql == 0) quad_st(&out->x, X3);
    }
    __syncwarp(mask);
    if (ql == 0) quad_st(&out->y, sub(r4, quad_ld(&sc->v[10])));
    __syncwarp(mask);
}

// ---- general form: q given as (x = *qx, y = +-q->y, zz, zzz of *q) -----------------------------------
// *out = *p + Q with Q = (*qx, negq ? -q->y : q->y, q->zz, q->zzz): lets a caller add -T or the
// endomorphism image (beta x, y) of a table entry T without first building it in memory.
static __device__ __noinline__ void g1_add_mem_q(G1* out, const G1* p, const Fp* qx, const G1* q, bool negq) {
    G1 a, b;
    a.x = quad_ld(&p->x); a.y = quad_ld(&p->y); a.zz = quad_ld(&p->zz); a.zzz = quad_ld(&p->zzz);
    b.x = quad_ld(qx); b.y = quad_ld(&q->y); b.zz = quad_ld(&q->zz); b.zzz = quad_ld(&q->zzz);
    if (negq) b.y = neg(b.y);
    g1_add_to(a, b);
    quad_st(&out->x, a.x); quad_st(&out->y, a.y); quad_st(&out->zz, a.zz); quad_st(&out->zzz, a.zzz);
}

static __device__ __noinline__ void g1_add_quad_q(G1* out, const G1* p, const Fp* qx, const G1* q, bool negq, QuadScratch* sc) {
    const unsigned lane = threadIdx.x & 31u, ql = lane & 3u;
    const unsigned mask = 0xFu << (lane & ~3u);
    const Fp pzz = quad_ld(&p->zz), qzz = quad_ld(&q->zz);
    const bool pinf = is_zero(pzz), qinf = is_zero(qzz);
    if (pinf || qinf) {
        __syncwarp(mask);
        if (ql == 0) {
            if (pinf) {
                Fp a = quad_ld(qx), b = quad_ld(&q->y), c = quad_ld(&q->zz), d = quad_ld(&q->zzz);
                if (negq) b = neg(b);
                quad_st(&out->x, a); quad_st(&out->y, b); quad_st(&out->zz, c); quad_st(&out->zzz, d);
            } else if (out != p) {
                Fp a = quad_ld(&p->x), b = quad_ld(&p->y), c = quad_ld(&p->zz), d = quad_ld(&p->zzz);
                quad_st(&out->x, a); quad_st(&out->y, b); quad_st(&out->zz, c); quad_st(&out->zzz, d);
            }
        }
        __syncwarp(mask);
        return;
    }
    // level 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
    {
        const Fp* ap = ql == 0 ? &p->x : ql == 1 ? qx : ql == 2 ? &p->y : &q->y;
        const Fp* bp = ql == 0 ? &q->zz : ql == 1 ? &p->zz : ql == 2 ? &q->zzz : &p->zzz;
        Fp a = quad_ld(ap);
        if (ql == 3 && negq) a = neg(a);
        quad_st(&sc->v[ql], mul(a, quad_ld(bp)));
    }
    __syncwarp(mask);
    const Fp Pd = sub(quad_ld(&sc->v[1]), quad_ld(&sc->v[0])), Rd = sub(quad_ld(&sc->v[3]), quad_ld(&sc->v[2]));
    if (is_zero(Pd)) {  // same x: doubling or cancellation -- rare, one lane does the complete addition
        __syncwarp(mask);
        if (ql == 0) g1_add_mem_q(out, p, qx, q, negq);
        __syncwarp(mask);
        return;
    }
    // level 2: PP = P^2, ZZ1 ZZ2, RR = R^2, ZZZ1 ZZZ2
    {
        const Fp* ap = (ql == 1) ? &p->zz : &p->zzz;
        const Fp* bp = (ql == 1) ? &q->zz : &q->zzz;
        const Fp la = quad_ld(ap), lb = quad_ld(bp);
        const Fp d = quad_sel(ql == 0, Pd, Rd);
        const bool even = (ql & 1u) == 0;
        quad_st(&sc->v[4 + ql], mul(quad_sel(even, d, la), quad_sel(even, d, lb)));
    }
    __syncwarp(mask);  // from here on the inputs are dead: `out` may be written even if it aliases them
    // level 3: PPP = P PP, Q = U1 PP, ZZ3 = (ZZ1 ZZ2) PP
    {
        const Fp a = quad_sel(ql == 1, quad_ld(&sc->v[0]), quad_sel(ql == 2, quad_ld(&sc->v[5]), Pd));
        const Fp r3 = mul(a, quad_ld(&sc->v[4]));
        if (ql == 0) quad_st(&sc->v[8], r3);
        if (ql == 1) quad_st(&sc->v[9], r3);
        if (ql == 2) quad_st(&out->zz, r3);
    }
    __syncwarp(mask);
    // level 4: X3 = RR - PPP - 2Q;  R (Q - X3),  S1 PPP,  ZZZ3 = (ZZZ1 ZZZ2) PPP
    const Fp PPP = quad_ld(&sc->v[8]), Qv = quad_ld(&sc->v[9]);
    const Fp X3 = sub(sub(quad_ld(&sc->v[6]), PPP), dbl(Qv));
    Fp r4;
    {
        const Fp a = quad_sel(ql == 1, quad_ld(&sc->v[2]), quad_sel(ql == 2, quad_ld(&sc->v[7]), Rd));
        const Fp b = quad_sel(ql == 0 || ql == 3, sub(Qv, X3), PPP);
        r4 = mul(a, b);
        if (ql == 1) quad_st(&sc->v[10], r4);
        if (ql == 2) quad_st(&out->zzz, r4);
        if (ql == 0) quad_st(&out->x, X3);
    }
    __syncwarp(mask);
    if (ql == 0) quad_st(&out->y, sub(r4, quad_ld(&sc->v[10])));
    __syncwarp(mask);
}

// *out = 2 * *p on the four lanes of a quad (dbl-2008-s-1 has depth three: {V = (2Y)^2, X^2} ->
// {W = 2Y V, S = X V, ZZ3 = V ZZ, M^2} -> {M (S - X3), W Y, ZZZ3 = W ZZZ}); `out` may alias `p`.
static __device__ __noinline__ void g1_dbl_quad(G1* out, const G1* p, QuadScratch* sc) {
    const unsigned lane = threadIdx.x & 31u, ql = lane & 3u;
    const unsigned mask = 0xFu << (lane & ~3u);
    const Fp pzz = quad_ld(&p->zz);
    if (is_zero(pzz)) {  // infinity stays infinity
        __syncwarp(mask);
        if (ql == 0 && out != p) {
            const Fp z = Fp::zero();
            quad_st(&out->x, z); quad_st(&out->y, z); quad_st(&out->zz, z); quad_st(&out->zzz, z);
        }
        __syncwarp(mask);
        return;
    }
    const Fp x = quad_ld(&p->x), y = quad_ld(&p->y);
    const Fp U = dbl(y);
    // level 1 (two distinct products; lanes 2, 3 repeat them)
    {
        const Fp a = quad_sel((ql & 1u) == 0, U, x);
        const Fp r1 = mul(a, a);
        if (ql < 2) quad_st(&sc->v[ql], r1);  // v[0] = V, v[1] = X^2
    }
    __syncwarp(mask);
    const Fp V = quad_ld(&sc->v[0]);
    Fp M = quad_ld(&sc->v[1]);
    M = add(dbl(M), M);
    // level 2
    Fp r2;
    {
        const Fp a = quad_sel(ql == 0, U, quad_sel(ql == 1, x, quad_sel(ql == 2, pzz, M)));
        const Fp b = quad_sel(ql == 3, M, V);
        r2 = mul(a, b);
        if (ql == 0) quad_st(&sc->v[2], r2);  // W
        if (ql == 1) quad_st(&sc->v[3], r2);  // S
        if (ql == 3) quad_st(&sc->v[4], r2);  // M^2
    }
    const Fp pzzz = quad_ld(&p->zzz);
    __syncwarp(mask);
    // level 3
    const Fp W = quad_ld(&sc->v[2]), S = quad_ld(&sc->v[3]);
    const Fp X3 = sub(quad_ld(&sc->v[4]), dbl(S));
    Fp r3;
    {
        const Fp a = quad_sel(ql == 0, M, W);
        const Fp b = quad_sel(ql == 0, sub(S, X3), quad_sel(ql == 1, y, pzzz));
        r3 = mul(a, b);
        if (ql == 1) quad_st(&sc->v[5], r3);  // W Y
    }
    __syncwarp(mask);  // every lane has read what it needs of *p
    if (ql == 2) {
        quad_st(&out->zz, r2);
        quad_st(&out->zzz, r3);
    }
    if (ql == 0) {
        quad_st(&out->x, X3);
        quad_st(&out->y, sub(r3, quad_ld(&sc->v[5])));
    }
    __syncwarp(mask);
}

// *dst = *src, the twelve 16-byte words spread over the quad; callers synchronise
__device__ __forceinline__ void quad_copy_g1(G1* dst, const G1* src) {
    const unsigned ql = threadIdx.x & 3u;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 3; i++) d[ql * 3 + i] = s[ql * 3 + i];
}

// ---- validation on a quad: decompression on lane 0, the two |z| chains of the subgroup test shared ------
// Same result as g1a_validate_levels (g1.cuh): decompress, accept infinity, else require phi(P) + [z^2]P = inf,
// and leave the 18 table levels (2^(8j) P, 2^(8j) [|z|]P) behind.  The chains are 2 x (63 doublings + 5
// additions): 1.2 ms of dependent products on one lane, 0.4 ms on a quad.
struct QuadValidate {
    G1 p;    // the point (zz = zzz = 1)
    G1 q;    // [|z|] P
    G1 d;    // running doubling chain
    G1 acc;
    Fp bx;   // beta * x
    QuadScratch sc;
    int state;  // 0 invalid encoding, 1 finite point, 2 infinity
    int pad[3];
};

// *acc = [|z|] *src, storing 2^(8j) *src (j = 0..8) at levels[j * stride]; *d is scratch
static __device__ __noinline__ void g1_mul_bls_x_levels_quad(G1* acc, G1* d, const G1* src, G1* levels, size_t stride, QuadScratch* sc) {
    const unsigned lane = threadIdx.x & 31u, ql = lane & 3u;
    const unsigned mask = 0xFu << (lane & ~3u);
    quad_copy_g1(d, src);
    if (ql == 0) {
        const Fp z = Fp::zero();
        quad_st(&acc->x, z); quad_st(&acc->y, z); quad_st(&acc->zz, z); quad_st(&acc->zzz, z);
    }
    __syncwarp(mask);
    const uint64_t x = BLS_X_ABS;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        if ((i & 7) == 0) {
            quad_copy_g1(levels + (size_t)(i >> 3) * stride, d);
            __syncwarp(mask);
        }
        if ((x >> i) & 1ull) g1_add_quad(acc, acc, d, sc);
        g1_dbl_quad(d, d, sc);
    }
    quad_copy_g1(levels + (size_t)8 * stride, d);
    __syncwarp(mask);
}

// Called by the four lanes of a quad with identical arguments; returns the verdict to all of them.
// out_affine (may be null) receives the decompressed point (infinity when invalid).
static __device__ __noinline__ bool g1a_validate_levels_quad(G1Affine* out_affine, const uint8_t* in48, G1* levels, size_t stride, QuadValidate* W) {
    const unsigned lane = threadIdx.x & 31u, ql = lane & 3u;
    const unsigned mask = 0xFu << (lane & ~3u);
    if (ql == 0) {
        uint8_t buf[48];
        for (int i = 0; i < 48; i++) buf[i] = in48[i];
        G1Affine a;
        const bool ok = g1a_uncompress(a, buf);
        const int st = !ok ? 0 : (g1a_is_inf(a) ? 2 : 1);
        if (st == 1) {
            const Fp one = Fp::one();
            quad_st(&W->p.x, a.x); quad_st(&W->p.y, a.y); quad_st(&W->p.zz, one); quad_st(&W->p.zzz, one);
            quad_st(&W->bx, mul(a.x, Fp::from_limbs(FP_BETA_A)));
        }
        if (out_affine) *out_affine = (st == 1) ? a : g1a_inf();
        W->state = st;
    }
    __syncwarp(mask);
    const int st = W->state;
    if (st != 1) {
        const G1 inf = g1_inf();
        for (int j = (int)ql; j < G1_LEVELS; j += 4) g1_store(levels + (size_t)j * stride, inf);
        __syncwarp(mask);
        return st == 2;
    }
    g1_mul_bls_x_levels_quad(&W->acc, &W->d, &W->p, levels, stride, &W->sc);
    quad_copy_g1(&W->q, &W->acc);
    __syncwarp(mask);
    g1_mul_bls_x_levels_quad(&W->acc, &W->d, &W->q, levels + (size_t)9 * stride, stride, &W->sc);
    g1_add_quad_q(&W->acc, &W->acc, &W->bx, &W->p, false, &W->sc);  // phi(P) + [z^2]P
    const bool in_group = is_zero(quad_ld(&W->acc.zz));
    __syncwarp(mask);
    return in_group;
}

}  // namespace kzg
