// Hot-loop form of the mixed addition: every product goes through ONE out-of-line multiplier that takes
// its operands by value (registers), so the loop body stays small enough for the instruction cache and
// nothing is forced into local memory.  Used by the bucket accumulation (msm.cu) and the FK20 table
// MSMs (fk20.cu).  ncu r01b: "no_instruction" stalls ~1 per issue with ten multipliers inlined.
#pragma once
#include "g1.cuh"

namespace kzg {

// out-of-line multiplier for the hot loop: keeps the loop body small enough for the instruction
// cache (ncu r01b: "no_instruction" stalls ~1 per issue with ten multipliers inlined)
static __device__ __noinline__ Fp fp_mul_nl(Fp a, Fp b) { return mul(a, b); }

// acc += +-a with every product through fp_mul_nl (same formulas as g1_madd in g1.cuh)
__device__ __forceinline__ void g1_madd_nl(G1& acc, const G1Affine& a_in, bool negate) {
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
    Fp U2 = fp_mul_nl(a.x, acc.zz);
    Fp S2 = fp_mul_nl(a.y, acc.zzz);
    Fp Pd = sub(U2, acc.x);
    Fp Rd = sub(S2, acc.y);
    if (is_zero(Pd)) {
        if (is_zero(Rd))
            acc = g1_dbl_affine(a);
        else
            acc = g1_inf();
        return;
    }
    Fp PP = fp_mul_nl(Pd, Pd);
    Fp PPP = fp_mul_nl(Pd, PP);
    Fp Q = fp_mul_nl(acc.x, PP);
    Fp X3 = sub(sub(fp_mul_nl(Rd, Rd), PPP), dbl(Q));
    Fp Y3 = sub(fp_mul_nl(Rd, sub(Q, X3)), fp_mul_nl(acc.y, PPP));
    acc.x = X3;
    acc.y = Y3;
    acc.zz = fp_mul_nl(acc.zz, PP);
    acc.zzz = fp_mul_nl(acc.zzz, PPP);
}


}  // namespace kzg
