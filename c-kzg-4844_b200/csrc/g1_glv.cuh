// Variable-scalar G1 multiplication for the linear combinations of the verifiers (device only).
//
// Replaces the per-point multiplications of g1_lincomb_naive / g1_mul (src/common/lincomb.c:34,
// src/common/ec.c:53 -> blst POINTonE1_mult_glv, blst/src/e1.c:412-442) and the small lincombs of
// verify_cell_kzg_proof_batch.
//
// GLV: k = k1 + k2*lambda, lambda = -z^2 (eigenvalue of phi(x,y) = (beta x, y)), obtained by plain
// division by z^2 (k1 = k mod z^2, k2 = -(k div z^2), both < 2^128), so [k]P = [k1]P + [q](-phi(P)).
// Signed 4-bit fixed windows over the two 128-bit halves share the doublings and ONE table of
// {1..8}P (the second base's multiples are phi of the first's): 33 x (4 dbl + 2 add) -- and the
// control flow is the same for every scalar, so lanes with different scalars do not diverge (the
// plain double-and-add executed ~255 additions per warp instead of the ~128 each lane needed).
#pragma once
#include "g1.cuh"

namespace kzg {

// z^2 for z = 0xd201000000010000 (little-endian 32-bit limbs)
__device__ __forceinline__ void glv_split(uint32_t k1[4], uint32_t q[4], const uint32_t k[8]) {
    const uint32_t Z2[5] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u, 0u};  // requires k < r (then k div z^2 < 2^128)
    uint32_t rem[5] = {0, 0, 0, 0, 0};
    q[0] = q[1] = q[2] = q[3] = 0;
#pragma unroll 1
    for (int bit = 255; bit >= 0; bit--) {
        // rem = (rem << 1) | k[bit]
        uint32_t in = (k[bit >> 5] >> (bit & 31)) & 1u;
#pragma unroll
        for (int i = 4; i > 0; i--) rem[i] = (rem[i] << 1) | (rem[i - 1] >> 31);
        rem[0] = (rem[0] << 1) | in;
        uint32_t t[5];
        uint32_t bw = limbs_sub<5>(t, rem, Z2);
        if (!bw) {
#pragma unroll
            for (int i = 0; i < 5; i++) rem[i] = t[i];
            if (bit < 128) q[bit >> 5] |= 1u << (bit & 31);
        }
    }
    k1[0] = rem[0]; k1[1] = rem[1]; k1[2] = rem[2]; k1[3] = rem[3];
}

// signed 4-bit digits of a 128-bit value: v = sum d[i] 16^i, d[i] in [-7, 8], i < 33
__device__ __forceinline__ void glv_digits(int8_t d[33], const uint32_t v[4]) {
    uint32_t carry = 0;
#pragma unroll 1
    for (int i = 0; i < 32; i++) {
        uint32_t x = ((v[i >> 3] >> ((i & 7) * 4)) & 0xfu) + carry;
        if (x > 8) {
            d[i] = (int8_t)((int)x - 16);
            carry = 1;
        } else {
            d[i] = (int8_t)x;
            carry = 0;
        }
    }
    d[32] = (int8_t)carry;
}

// [k]P, P given in XYZZ (not infinity-checked by the caller: infinity in -> infinity out)
static __device__ __noinline__ G1 g1_mul_glv(const G1& p, const uint32_t* k) {
    if (g1_is_inf(p)) return g1_inf();
    uint32_t k1[4], q[4];
    glv_split(k1, q, k);
    int8_t d1[33], d2[33];
    glv_digits(d1, k1);
    glv_digits(d2, q);
    G1 tab[8];  // (i+1) P
    tab[0] = p;
#pragma unroll 1
    for (int i = 1; i < 8; i++) {
        tab[i] = tab[i - 1];
        g1_add_to(tab[i], p);  // i == 1 hits the doubling branch of the complete addition
    }
    const Fp beta = Fp::from_limbs(FP_BETA_A);
    G1 acc = g1_inf();
#pragma unroll 1
    for (int i = 32; i >= 0; i--) {
#pragma unroll 1
        for (int s = 0; s < 4; s++) g1_dbl_to(acc);
        int a = d1[i], b = d2[i];
        if (a != 0) {
            G1 t = tab[(a < 0 ? -a : a) - 1];
            if (a < 0) t.y = neg(t.y);
            g1_add_to(acc, t);
        }
        if (b != 0) {
            G1 t = tab[(b < 0 ? -b : b) - 1];
            t.x = mul(t.x, beta);
            if (b > 0) t.y = neg(t.y);  // second base is -phi(P)
            g1_add_to(acc, t);
        }
    }
    return acc;
}

__device__ __forceinline__ G1 g1_mul_glv_affine(const G1Affine& a, const uint32_t* k) { return g1_mul_glv(g1_from_affine(a), k); }

}  // namespace kzg
