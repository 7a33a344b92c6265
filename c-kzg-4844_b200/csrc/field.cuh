// Montgomery prime-field arithmetic for BLS12-381 on 32-bit limbs (Fp: 12 limbs, Fr: 8 limbs).
//
// Replaces, for the GPU engine, what the reference gets from blst's assembly:
//   mul_mont_384 / mul_mont_sparse_256, add_mod_*, sub_mod_*, cneg_mod_*   (blst/src/vect.h:88-140,
//   blst/src/asm/mulx_mont_{384,256}-x86_64.pl), field shortcuts (blst/src/fields.h:14-50) and the
//   blst_fr_* exports (blst/src/exports.c:24-110).
//
// Device path (nvcc): PTX carry chains. One row of the product is split into the even-indexed and the
// odd-indexed limbs of `a`; each half is a single mad.lo.cc/madc.hi.cc chain over disjoint 64-bit
// slots, which ptxas fuses into IMAD.WIDE.U32(.X) -- ~N^2 wide MADs for the product and ~N^2 for the
// interleaved Montgomery reduction (checked with cuobjdump: 276 IMAD.WIDE per Fp mul).
// Host path (g++): the same functions in portable C on uint64_t. It exists ONLY so that
// tests/hostcheck can unit-test the formulas built on top of this header against the Python oracle
// in the GPU-less build container; it is never compiled into the product library.
//
// All values are kept fully reduced (< modulus) so equality is limb equality.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KZG_HD __device__ __forceinline__
#define KZG_HDS static __device__ __forceinline__
#define KZG_HD_NOINLINE static __device__ __noinline__
#define KZG_CONST static __device__ __constant__ const
#define KZG_DEVICE_PATH 1
#else
#define KZG_HD static inline
#define KZG_HDS static inline
#define KZG_HD_NOINLINE static
#define KZG_CONST static const
#define KZG_DEVICE_PATH 0
#endif

#include "constants.cuh"

namespace kzg {

template <int N>
struct Limbs {
    uint32_t l[N];
};

struct FpTag {
    static constexpr int N = 12;
    static constexpr uint32_t INV = FP_INV32;
    KZG_HDS const uint32_t* mod() { return FP_MOD; }
    KZG_HDS const uint32_t* one() { return FP_ONE; }
    KZG_HDS const uint32_t* r2() { return FP_R2; }
};
struct FrTag {
    static constexpr int N = 8;
    static constexpr uint32_t INV = FR_INV32;
    KZG_HDS const uint32_t* mod() { return FR_MOD; }
    KZG_HDS const uint32_t* one() { return FR_ONE; }
    KZG_HDS const uint32_t* r2() { return FR_R2; }
};

// ------------------------------------------------------------------------------------------------
// raw multi-limb helpers
// ------------------------------------------------------------------------------------------------

// r = a + b, returns carry
template <int N>
KZG_HD uint32_t limbs_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if KZG_DEVICE_PATH
    uint32_t c;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
    for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(a[i]), "r"(b[i]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < N; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}

// r = a - b, returns borrow (1 if a < b)
template <int N>
KZG_HD uint32_t limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if KZG_DEVICE_PATH
    uint32_t bw;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
    for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(a[i]), "r"(b[i]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(bw));
    return bw & 1u;
#else
    uint64_t bw = 0;
    for (int i = 0; i < N; i++) {
        uint64_t v = (uint64_t)a[i] - b[i] - bw;
        r[i] = (uint32_t)v;
        bw = (v >> 32) & 1;
    }
    return (uint32_t)bw;
#endif
}

// acc = acc + v (m == 0) or acc - v (m == 0xffffffff), modulo 2^(32N): one carry chain over v ^ m with
// the carry-in taken from m, so lanes that add and lanes that subtract run the same instructions
template <int N>
KZG_HD void limbs_addsub(uint32_t* acc, const uint32_t* v, uint32_t m) {
#if KZG_DEVICE_PATH
    uint32_t t;
    asm volatile("add.cc.u32 %0, %1, %1;" : "=r"(t) : "r"(m));  // CF = (m != 0)
#pragma unroll
    for (int i = 0; i < N - 1; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(v[i] ^ m));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(acc[N - 1]) : "r"(v[N - 1] ^ m));
#else
    uint64_t c = m & 1u;
    for (int i = 0; i < N; i++) {
        c += (uint64_t)acc[i] + (v[i] ^ m);
        acc[i] = (uint32_t)c;
        c >>= 32;
    }
#endif
}

// a >= b ?
template <int N>
KZG_HD bool limbs_geq(const uint32_t* a, const uint32_t* b) {
    uint32_t t[N];
    return limbs_sub<N>(t, a, b) == 0;
}

template <int N>
KZG_HD bool limbs_is_zero(const uint32_t* a) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc |= a[i];
    return acc == 0;
}

template <int N>
KZG_HD bool limbs_eq(const uint32_t* a, const uint32_t* b) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc |= a[i] ^ b[i];
    return acc == 0;
}

// ------------------------------------------------------------------------------------------------
// Montgomery multiplication
// ------------------------------------------------------------------------------------------------
#if KZG_DEVICE_PATH
namespace detail {
// acc[j], acc[j+1] = a[j] * w  for j = 0,2,4,... (n must be even)
template <int n>
__device__ __forceinline__ void row_mul(uint32_t* acc, const uint32_t* a, uint32_t w) {
#pragma unroll
    for (int j = 0; j < n; j += 2)
        asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(w));
}
// acc[0..n) += a[0,2,4,...] * w as one carry chain; the carry out is left in CC.CF
template <int n>
__device__ __forceinline__ void row_mad(uint32_t* acc, const uint32_t* a, uint32_t w) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(w));
#pragma unroll
    for (int j = 2; j < n; j += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(w));
}
// acc[j] = a[j]*w + acc[j+2] (the accumulator slides down two limbs), consuming the incoming CC.CF
template <int n>
__device__ __forceinline__ void row_mad_slide(uint32_t* acc, const uint32_t* a, uint32_t w) {
#pragma unroll
    for (int j = 0; j < n - 2; j += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
            : "=r"(acc[j]), "=r"(acc[j + 1])
            : "r"(a[j]), "r"(w), "r"(acc[j + 2]), "r"(acc[j + 3]));
    asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(acc[n - 2]), "=r"(acc[n - 1]) : "r"(a[n - 2]), "r"(w));
}
// One row of the interleaved product + reduction.  Invariant between rows:
//   T = sum lo[k] W^k + sum hi[k] W^(k+1)      (W = 2^32; `lo` holds the slots starting at even limb
//   positions, `hi` those starting at odd positions).  After adding a*w and m*p limb 0 is zero; the
//   division by W swaps the roles of the two arrays, which is why callers alternate (lo,hi).
template <class F, int n>
__device__ __forceinline__ void mont_row(uint32_t* lo, uint32_t* hi, const uint32_t* a, uint32_t w, bool first) {
    if (first) {
        row_mul<n>(hi, a + 1, w);
        row_mul<n>(lo, a, w);
    } else {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo[0]) : "r"(hi[1]));
        row_mad_slide<n>(hi, a + 1, w);
        row_mad<n>(lo, a, w);
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(hi[n - 1]));
    }
    uint32_t m = lo[0] * F::INV;
    row_mad<n>(hi, F::mod() + 1, m);
    row_mad<n>(lo, F::mod(), m);
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(hi[n - 1]));
}
}  // namespace detail
#endif

// r = a * b / R mod p, inputs and output fully reduced
template <class F>
KZG_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    constexpr int N = F::N;
#if KZG_DEVICE_PATH
    uint32_t ev[N], od[N];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        detail::mont_row<F, N>(ev, od, a, b[i], i == 0);
        detail::mont_row<F, N>(od, ev, a, b[i + 1], false);
    }
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(ev[0]) : "r"(od[1]));
#pragma unroll
    for (int i = 1; i < N - 1; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(ev[i]) : "r"(od[i + 1]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[N - 1]));
    uint32_t s[N];
    uint32_t bw = limbs_sub<N>(s, ev, F::mod());
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = bw ? ev[i] : s[i];
#else
    uint32_t t[N + 2];
    for (int i = 0; i < N + 2; i++) t[i] = 0;
    for (int i = 0; i < N; i++) {
        uint64_t c = 0;
        for (int j = 0; j < N; j++) {
            uint64_t v = (uint64_t)a[j] * b[i] + t[j] + c;
            t[j] = (uint32_t)v;
            c = v >> 32;
        }
        uint64_t v = (uint64_t)t[N] + c;
        t[N] = (uint32_t)v;
        t[N + 1] = (uint32_t)(v >> 32);
        uint32_t m = t[0] * F::INV;
        v = (uint64_t)m * F::mod()[0] + t[0];
        c = v >> 32;
        for (int j = 1; j < N; j++) {
            v = (uint64_t)m * F::mod()[j] + t[j] + c;
            t[j - 1] = (uint32_t)v;
            c = v >> 32;
        }
        v = (uint64_t)t[N] + c;
        t[N - 1] = (uint32_t)v;
        t[N] = t[N + 1] + (uint32_t)(v >> 32);
    }
    uint32_t s[N];
    uint32_t bw = limbs_sub<N>(s, t, F::mod());
    bool ge = t[N] != 0 || bw == 0;
    for (int i = 0; i < N; i++) r[i] = ge ? s[i] : t[i];
#endif
}

// ------------------------------------------------------------------------------------------------
// Field element wrapper
// ------------------------------------------------------------------------------------------------
template <class F>
struct Fe {
    static constexpr int N = F::N;
    uint32_t l[N];

    KZG_HDS Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    KZG_HDS Fe one() {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = F::one()[i];
        return r;
    }
    KZG_HDS Fe from_limbs(const uint32_t* p) {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = p[i];
        return r;
    }
};

template <class F>
KZG_HD bool is_zero(const Fe<F>& a) {
    return limbs_is_zero<F::N>(a.l);
}
template <class F>
KZG_HD bool eq(const Fe<F>& a, const Fe<F>& b) {
    return limbs_eq<F::N>(a.l, b.l);
}

template <class F>
KZG_HD Fe<F> add(const Fe<F>& a, const Fe<F>& b) {
    constexpr int N = F::N;
    Fe<F> r;
    uint32_t t[N], s[N];
    uint32_t c = limbs_add<N>(t, a.l, b.l);  // both moduli leave the top bit clear => c == 0
    (void)c;
    uint32_t bw = limbs_sub<N>(s, t, F::mod());
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = bw ? t[i] : s[i];
    return r;
}

template <class F>
KZG_HD Fe<F> sub(const Fe<F>& a, const Fe<F>& b) {
    constexpr int N = F::N;
    Fe<F> r;
    uint32_t t[N], s[N];
    uint32_t bw = limbs_sub<N>(t, a.l, b.l);
    limbs_add<N>(s, t, F::mod());
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = bw ? s[i] : t[i];
    return r;
}

template <class F>
KZG_HD Fe<F> neg(const Fe<F>& a) {
    constexpr int N = F::N;
    Fe<F> r;
    uint32_t t[N];
    limbs_sub<N>(t, F::mod(), a.l);
    bool z = limbs_is_zero<N>(a.l);
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = z ? 0u : t[i];
    return r;
}

template <class F>
KZG_HD Fe<F> cneg(const Fe<F>& a, bool flag) {
    Fe<F> n = neg(a), r;
#pragma unroll
    for (int i = 0; i < F::N; i++) r.l[i] = flag ? n.l[i] : a.l[i];
    return r;
}

template <class F>
KZG_HD Fe<F> select(bool flag, const Fe<F>& a, const Fe<F>& b) {  // flag ? a : b
    Fe<F> r;
#pragma unroll
    for (int i = 0; i < F::N; i++) r.l[i] = flag ? a.l[i] : b.l[i];
    return r;
}

template <class F>
KZG_HD Fe<F> dbl(const Fe<F>& a) {
    return add(a, a);
}

// KZG_FP_MUL_OUTLINE (defined by a .cu file before its includes): every 12-limb product of that
// translation unit becomes a call to ONE out-of-line multiplier taking its operands in registers.
// The kernels built from dependent chains of Fp products (point validation, scalar multiplication,
// tree sums) are then a few tens of KB of code instead of 150-450 KB (one inlined multiplier is 6.4 KB
// of SASS).  Their speed alone barely changes, but kernels that share an SM no longer evict each
// other's instructions: measured r01r, hash (39 KB loop) beside validation (404 KB): 2.06 ms -> 3.7-5 ms.
#if KZG_DEVICE_PATH && defined(KZG_FP_MUL_OUTLINE)
template <class F>
static __device__ __noinline__ Fe<F> fe_mul_outlined(Fe<F> a, Fe<F> b) {
    Fe<F> r;
    mont_mul<F>(r.l, a.l, b.l);
    return r;
}
#endif

template <class F>
KZG_HD Fe<F> mul(const Fe<F>& a, const Fe<F>& b) {
#if KZG_DEVICE_PATH && defined(KZG_FP_MUL_OUTLINE)
    if constexpr (F::N == 12) return fe_mul_outlined<F>(a, b);
#endif
    Fe<F> r;
    mont_mul<F>(r.l, a.l, b.l);
    return r;
}

template <class F>
KZG_HD Fe<F> sqr(const Fe<F>& a) {
#if KZG_DEVICE_PATH && defined(KZG_FP_MUL_OUTLINE)
    if constexpr (F::N == 12) return fe_mul_outlined<F>(a, a);
#endif
    Fe<F> r;
    mont_mul<F>(r.l, a.l, a.l);
    return r;
}

// plain integer (< modulus) -> Montgomery form
template <class F>
KZG_HD Fe<F> to_mont(const uint32_t* plain) {
    Fe<F> r;
    mont_mul<F>(r.l, plain, F::r2());
    return r;
}

// Montgomery form -> plain integer limbs
template <class F>
KZG_HD void from_mont(uint32_t* plain, const Fe<F>& a) {
    uint32_t one[F::N];
#pragma unroll
    for (int i = 0; i < F::N; i++) one[i] = (i == 0);
    mont_mul<F>(plain, a.l, one);
}

// a^e, e given as NL plain little-endian limbs (public exponent; variable time is fine: nothing here
// is secret).  MSB-first square-and-multiply.  Kept out of line: it is called from cold paths
// (inversion, square roots) and inlining it everywhere would bloat the hot kernels.
template <class F, int NL>
KZG_HD_NOINLINE Fe<F> pow_limbs(const Fe<F>& a, const uint32_t* e) {
    Fe<F> r = Fe<F>::one();
    bool started = false;
    for (int i = NL - 1; i >= 0; i--) {
        uint32_t w = e[i];
        for (int b = 31; b >= 0; b--) {
            if (started) r = sqr(r);
            if ((w >> b) & 1u) {
                r = started ? mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}

using Fp = Fe<FpTag>;
using Fr = Fe<FrTag>;

// Inversion by the binary extended Euclid in Kaliski's "almost Montgomery inverse" form (IEEE Trans.
// Computers 44(8), 1995): phase 1 uses only shifts, additions and subtractions and returns
// a^-1 * 2^k (mod p) with n <= k <= 2n; the power of two is then folded away with four Montgomery
// multiplications.  ~55 simple instructions per step and ~1.4 n steps -- roughly 7x fewer
// instructions and a much shorter dependent chain than the Fermat power (570 dependent Montgomery
// products for Fp), which is what the latency-bound single-thread callers care about (affine
// conversion before compression / the pairing, the Fp2 inverse of the final exponentiation, the
// per-blob barycentric denominator).  Variable time: every input here is public.
// Replaces what the reference gets from blst's ct_inverse_mod_383/256 (blst/src/recip.c:18-75).
// 0 -> 0, like the Fermat form it replaces.
template <int N>
KZG_HD void limbs_shr1(uint32_t* a) {
#pragma unroll
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] >>= 1;
}
template <int N>
KZG_HD void limbs_shl1(uint32_t* a) {
#pragma unroll
    for (int i = N - 1; i > 0; i--) a[i] = (a[i] << 1) | (a[i - 1] >> 31);
    a[0] <<= 1;
}
// shifts by t bits, 1 <= t <= 31 (funnel shifts: the same instruction count as a shift by one)
template <int N>
KZG_HD void limbs_shr(uint32_t* a, int t) {
#pragma unroll
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> t) | (a[i + 1] << (32 - t));
    a[N - 1] >>= t;
}
template <int N>
KZG_HD void limbs_shl(uint32_t* a, int t) {
#pragma unroll
    for (int i = N - 1; i > 0; i--) a[i] = (a[i] << t) | (a[i - 1] >> (32 - t));
    a[0] <<= t;
}
// trailing zero bits of a nonzero word
KZG_HD int word_ctz(uint32_t w) {
#if KZG_DEVICE_PATH
    return __ffs((int)w) - 1;
#else
    return __builtin_ctz(w);
#endif
}

// Steps of Kaliski's phase 1 that only shift are merged: after a subtraction of two odd numbers the difference has
// two trailing zero bits on average, and "v >>= 1, r <<= 1, k++" repeated t times is "v >>= t, r <<= t, k += t" -- the
// same sequence of values, ~0.7 n iterations instead of ~1.4 n (the subtraction step brings its own first halving).
template <class F>
KZG_HD_NOINLINE Fe<F> inv_binary(const Fe<F>& a) {
    constexpr int N = F::N;
    if (limbs_is_zero<N>(a.l)) return a;
    uint32_t u[N], v[N], r[N], s[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        u[i] = F::mod()[i];
        v[i] = a.l[i];
        r[i] = 0;
        s[i] = (i == 0);
    }
    int k = 0;
    // invariants: u, v > 0 until the end; gcd(u, v) = 1; r, s < 2p (fits: 2p < 2^(32N))
    while (!limbs_is_zero<N>(v)) {
        if (!(v[0] & 1u)) {
            const int t = v[0] ? word_ctz(v[0]) : 31;  // a zero low word: 31 now, the rest in the next rounds
            limbs_shr<N>(v, t);
            limbs_shl<N>(r, t);
            k += t;
        } else if (!(u[0] & 1u)) {
            const int t = u[0] ? word_ctz(u[0]) : 31;
            limbs_shr<N>(u, t);
            limbs_shl<N>(s, t);
            k += t;
        } else {
            uint32_t d[N];
            if (limbs_sub<N>(d, v, u) == 0) {  // v >= u
                // v - u is even; if it is zero the loop ends after this one halving (u = v = gcd = 1)
                const int t = limbs_is_zero<N>(d) ? 1 : (d[0] ? word_ctz(d[0]) : 31);
#pragma unroll
                for (int i = 0; i < N; i++) v[i] = d[i];
                limbs_shr<N>(v, t);
                limbs_add<N>(s, s, r);
                limbs_shl<N>(r, t);
                k += t;
            } else {
                limbs_sub<N>(u, u, v);  // u > v: the difference is even and not zero
                const int t = u[0] ? word_ctz(u[0]) : 31;
                limbs_shr<N>(u, t);
                limbs_add<N>(r, r, s);
                limbs_shl<N>(s, t);
                k += t;
            }
        }
    }
    if (limbs_geq<N>(r, F::mod())) limbs_sub<N>(r, r, F::mod());
    Fe<F> x;  // x = p - r = (integer a)^-1 * 2^k  (mod p), as a plain integer
    limbs_sub<N>(x.l, F::mod(), r);
    // The input integer is a_true * R, so the Montgomery form of the inverse is x * 2^(2*32N - k).
    // Multiply by that power of two in two pieces that are each < p (2^(32N-4) < p for both fields).
    int m = 2 * 32 * N - k;
    int m1 = m < 32 * N - 4 ? m : 32 * N - 4;
    int m2 = m - m1;
    Fe<F> t1, t2, r2;
#pragma unroll
    for (int i = 0; i < N; i++) {
        t1.l[i] = (i == (m1 >> 5)) ? (1u << (m1 & 31)) : 0u;
        t2.l[i] = (i == (m2 >> 5)) ? (1u << (m2 & 31)) : 0u;
        r2.l[i] = F::r2()[i];
    }
    x = mul(mul(x, t1), r2);  // x * 2^m1
    x = mul(mul(x, t2), r2);  // x * 2^m
    return x;
}

KZG_HD Fp fp_inv(const Fp& a) { return inv_binary<FpTag>(a); }  // 0 -> 0
KZG_HD Fr fr_inv(const Fr& a) { return inv_binary<FrTag>(a); }  // 0 -> 0
// the Fermat forms stay for the self-test, which cross-checks the two on the device
KZG_HD Fp fp_inv_fermat(const Fp& a) { return pow_limbs<FpTag, 12>(a, FP_P_MINUS_2); }
KZG_HD Fr fr_inv_fermat(const Fr& a) { return pow_limbs<FrTag, 8>(a, FR_R_MINUS_2); }

// ------------------------------------------------------------------------------------------------
// big-endian byte <-> limb conversion (wire format: src/common/bytes.c:52-70)
// ------------------------------------------------------------------------------------------------
template <int N>
KZG_HD void limbs_from_be(uint32_t* l, const uint8_t* b) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint8_t* q = b + 4 * (N - 1 - i);
        l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
}
template <int N>
KZG_HD void limbs_to_be(uint8_t* b, const uint32_t* l) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint8_t* q = b + 4 * (N - 1 - i);
        q[0] = (uint8_t)(l[i] >> 24);
        q[1] = (uint8_t)(l[i] >> 16);
        q[2] = (uint8_t)(l[i] >> 8);
        q[3] = (uint8_t)l[i];
    }
}

// bytes_to_bls_field (src/common/bytes.c:64): false if the big-endian value is >= r
KZG_HD bool fr_from_be_checked(Fr& out, const uint8_t* b) {
    uint32_t t[8];
    limbs_from_be<8>(t, b);
    bool ok = !limbs_geq<8>(t, FR_MOD);
    out = to_mont<FrTag>(t);
    return ok;
}
// bytes_from_bls_field (src/common/bytes.c:52)
KZG_HD void fr_to_be(uint8_t* b, const Fr& a) {
    uint32_t t[8];
    from_mont<FrTag>(t, a);
    limbs_to_be<8>(b, t);
}
// hash_to_bls_field (src/common/bytes.c:123): 256-bit big-endian value reduced mod r (2^256 < 3r)
KZG_HD Fr fr_from_be_reduce(const uint8_t* b) {
    uint32_t t[8], s[8];
    limbs_from_be<8>(t, b);
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        uint32_t bw = limbs_sub<8>(s, t, FR_MOD);
        if (!bw) {
            for (int i = 0; i < 8; i++) t[i] = s[i];
        }
    }
    return to_mont<FrTag>(t);
}

KZG_HD Fr fr_from_u64(uint64_t v) {
    uint32_t t[8] = {(uint32_t)v, (uint32_t)(v >> 32), 0, 0, 0, 0, 0, 0};
    return to_mont<FrTag>(t);
}

}  // namespace kzg
