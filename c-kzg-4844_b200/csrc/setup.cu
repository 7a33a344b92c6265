// Device-side trusted-setup preparation (replaces the loops of load_trusted_setup,
// src/setup/setup.c:392-505): batched G1 decompression, bit-reversal of the Lagrange points
// (setup.c:488), roots of unity (setup.c:99-153) and the fixed-base window table of msm.cu.
#include "engine.h"

namespace kzg {

__device__ __forceinline__ uint32_t bit_reverse(uint32_t v, int bits) { return __brev(v) >> (32 - bits); }

// out[i] = uncompress(bytes + 48 * src(i)), src(i) = bit-reversed i when requested.
// blst_p1_uncompress semantics (blst/src/e1.c:236-294); optional subgroup check = validate_kzg_g1
// (src/common/bytes.c:81).  Any failure sets *bad.
__global__ void g1_uncompress_kernel(G1Affine* __restrict__ out, const uint8_t* __restrict__ bytes, int n, int log2n_for_brp, bool check_subgroup, int* __restrict__ bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int src = log2n_for_brp ? (int)bit_reverse((uint32_t)i, log2n_for_brp) : i;
    uint8_t buf[48];
    for (int k = 0; k < 48; k++) buf[k] = bytes[(size_t)src * 48 + k];
    G1Affine a;
    bool ok = g1a_uncompress(a, buf);
    if (ok && check_subgroup && !g1a_is_inf(a)) ok = g1a_in_subgroup(a);
    if (!ok) {
        *bad = 1;
        a = g1a_inf();
    }
    out[i] = a;
}

// table[j][i] = 2^(MSM_C * j) * base[i], affine.  One thread per base point.
__global__ void msm_table_kernel(G1Affine* __restrict__ table, const G1Affine* __restrict__ base) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N_BLOB) return;
    G1Affine a = base[i];
    table[i] = a;
    G1 p = g1_from_affine(a);
#pragma unroll 1
    for (int j = 1; j < MSM_W; j++) {
#pragma unroll 1
        for (int k = 0; k < MSM_C; k++) g1_dbl_to(p);
        G1Affine q = g1_to_affine(p);
        table[(size_t)j * N_BLOB + i] = q;
        p = g1_from_affine(q);  // renormalise: keeps the doublings on zz = 1 inputs
    }
}

// roots[i] = w^i for i = 0..8192 (w = 7^((r-1)/8192)), roots_brp = brp(roots[0..8191]).
// One thread per index: w^i by square-and-multiply (13 bits) -- setup-time only.
__global__ void roots_kernel(Fr* __restrict__ roots, Fr* __restrict__ roots_brp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N_EXT) return;
    Fr w = Fr::from_limbs(FR_ROOT_8192);
    Fr acc = Fr::one();
    for (int b = 13; b >= 0; b--) {
        acc = sqr(acc);
        if ((i >> b) & 1) acc = mul(acc, w);
    }
    roots[i] = acc;
    if (i < N_EXT) roots_brp[bit_reverse((uint32_t)i, 13)] = acc;
}

int launch_g1_uncompress(Launch& L, G1Affine* out, const uint8_t* bytes, int n, bool brp, bool check_subgroup, int* d_bad) {
    if (n == 0) return RET_OK;
    int log2n = 0;
    if (brp) {
        while ((1 << log2n) < n) log2n++;
    }
    g1_uncompress_kernel<<<(n + 63) / 64, 64, 0, L.stream>>>(out, bytes, n, log2n, check_subgroup, d_bad);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}

int launch_msm_table(Launch& L, G1Affine* table, const G1Affine* base) {
    msm_table_kernel<<<N_BLOB / 64, 64, 0, L.stream>>>(table, base);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}

int launch_roots(Launch& L, Fr* roots, Fr* roots_brp) {
    roots_kernel<<<(N_EXT + 1 + 127) / 128, 128, 0, L.stream>>>(roots, roots_brp);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count();
    return RET_OK;
}

}  // namespace kzg
