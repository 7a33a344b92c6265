// Pairing-side kernels: G2 setup (decompression + line tables for the three fixed G2 arguments),
// the monomial-form sanity check of load_trusted_setup (src/setup/setup.c:339-358) and the final
// pairing check of the verifiers (pairings_verify, src/common/utils.c:172-196).
//
// One pairing check per verify call is a latency problem (~19k dependent Fp products), not a
// throughput one: it runs as ONE 64-thread CTA with the products of each Fp12
// operation spread over the lanes (pairing_coop.cuh).
#include "pairing.cuh"
#include "pairing_coop.cuh"
#include "verify.h"

namespace kzg {

__global__ void g2_uncompress_kernel(G2Affine* out, const uint8_t* bytes, int n, int* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[96];
    for (int k = 0; k < 96; k++) buf[k] = bytes[96 * i + k];
    G2Affine q;
    if (!g2a_uncompress(q, buf)) *bad = 1;
    out[i] = q;
}

// lines[0] <- G2[0] (generator slot), lines[1] <- G2[1] = [tau]G2, lines[2] <- G2[64] = [tau^64]G2
__global__ void g2_lines_kernel(G2Lines* lines, const G2Affine* g2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3) return;
    const int src[3] = {0, 1, 64};
    G2Affine q = g2[src[i]];
    g2_precompute_lines(lines[i], q);
}

// e(L[1], G2[0]) == e(L[0], G2[1])  <=>  e(-L[1], G2[0]) * e(L[0], G2[1]) == 1
// One CTA of COOP_LANES threads (pairing_coop.cuh); every thread reaches every barrier.
// The workspace (lane schedules, registers, all 2 x 68 evaluated lines: ~55 KB) is dynamic shared memory.
struct PairingSmem {
    CoopWS ws;
    G1 pts[2];
    int dec_ok[2];
};
// pairing_check_kernel: two cooperative machines (one per pairing of the product)
struct PairingSmem2 {
    CoopWS ws[2];
    G1 pts[2];
};
extern __shared__ __align__(16) unsigned char pairing_smem_raw[];

__global__ void __launch_bounds__(COOP_LANES) monomial_form_kernel(int* out, const uint8_t* lag01, const G2Lines* lines) {
    PairingSmem& sm = *reinterpret_cast<PairingSmem*>(pairing_smem_raw);
    CoopWS& ws = sm.ws;
    G1* pts = sm.pts;
    int* dec_ok = sm.dec_ok;
    if (threadIdx.x < 2) {
        uint8_t buf[48];
        G1Affine a;
        for (int k = 0; k < 48; k++) buf[k] = lag01[48 * threadIdx.x + k];
        dec_ok[threadIdx.x] = g1a_uncompress(a, buf) ? 1 : 0;
        pts[threadIdx.x] = g1_from_affine(a);
    }
    __syncthreads();
    // pts[0] = L[0], pts[1] = L[1]: first pairing argument is -L[1] against G2[0], second L[0] against G2[1]
    coop_pairing_product_is_one(ws, pts[1], &lines[0], pts[0], &lines[1], true);
    if (threadIdx.x == 0) *out = (dec_ok[0] && dec_ok[1]) ? ws.result : 0;  // decoding errors are reported by the main pass
}

// ok = [ e(A, Q_a) == e(B + B_extra, Q_b) ] = [ e(-A, Q_a) * e(B + B_extra, Q_b) == 1 ];  A, B XYZZ sums.
// 128 threads = TWO cooperative machines: the Miller loops of the two pairings are independent chains (63 squarings +
// 68 line products each), so they run side by side on the two halves of the CTA -- 131 dependent tower operations
// instead of the 199 of the shared-squaring loop on one machine -- and machine 0 multiplies the two values and runs
// the final exponentiation (the Miller value of a product of pairings is the product of the Miller values).
__global__ void __launch_bounds__(2 * COOP_LANES) pairing_check_kernel(int* ok, const G1* A, const G1* B, const G1* B_extra, const G2Lines* lines, int line_a, int line_b) {
    PairingSmem2& sm = *reinterpret_cast<PairingSmem2*>(pairing_smem_raw);
    const int g = threadIdx.x >> 6;
    CoopWS& ws = sm.ws[g];
    G1* pts = sm.pts;
    if (threadIdx.x == 0) {
        pts[0] = *A;
        G1 b = *B;
        if (B_extra) {
            G1 e = *B_extra;
            g1_add_to(b, e);
        }
        pts[1] = b;
    }
    __syncthreads();
    coop_init_tables(ws);
    coop_load_points(ws, pts[0], &lines[line_a], pts[1], &lines[line_b], true);
    if ((threadIdx.x & (COOP_LANES - 1)) == 0) ws.use[1 - g] = 0;  // machine g owns pairing g
    coop_sync();
    coop_prepare_all_lines(ws, &lines[line_a], &lines[line_b]);
    coop_miller_loop(ws);
    __syncthreads();
    if (g != 0) return;
    // machine 0: F = F_0 * F_1, final exponentiation
    const int lane = threadIdx.x;
    if (lane < 12) {
        ws.reg[6][lane] = sm.ws[1].reg[0][lane];
        ws.nreg[6][lane] = sm.ws[1].nreg[0][lane];
    }
    coop_sync();
    coop_mul(ws, 0, 0, 6);
    coop_final_exp_is_one(ws);
    if (threadIdx.x == 0) *ok = ws.result;
}

// > 48 KB of dynamic shared memory needs the per-function opt-in, once per device context
static int pairing_smem_opt_in() {
    KZG_CUDA_TRY(cudaFuncSetAttribute(monomial_form_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairingSmem)));
    KZG_CUDA_TRY(cudaFuncSetAttribute(pairing_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairingSmem2)));
    return RET_OK;
}

int setup_g2_and_lines(cudaStream_t stream, Launch& L, Ctx* c, const uint8_t* g2_host, int* d_bad) {
    int rc0 = pairing_smem_opt_in();
    if (rc0) return rc0;
    KZG_CUDA_TRY(cudaMalloc(&c->g2_points, 65 * sizeof(G2Affine)));
    KZG_CUDA_TRY(cudaMalloc(&c->g2_lines, 3 * sizeof(G2Lines)));
    uint8_t* d_bytes = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&d_bytes, 65 * 96, stream));
    KZG_CUDA_TRY(cudaMemcpyAsync(d_bytes, g2_host, 65 * 96, cudaMemcpyHostToDevice, stream));
    g2_uncompress_kernel<<<3, 32, 0, stream>>>((G2Affine*)c->g2_points, d_bytes, 65, d_bad);
    KZG_CUDA_TRY(cudaGetLastError());
    g2_lines_kernel<<<1, 32, 0, stream>>>((G2Lines*)c->g2_lines, (const G2Affine*)c->g2_points);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaFreeAsync(d_bytes, stream));
    L.count(2);
    return RET_OK;
}

int setup_is_monomial_form(cudaStream_t stream, Launch& L, Ctx* c, const uint8_t* g1_lagrange_host, int* is_monomial) {
    uint8_t* d_bytes = nullptr;
    int* d_out = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&d_bytes, 96, stream));
    KZG_CUDA_TRY(cudaMallocAsync((void**)&d_out, sizeof(int), stream));
    KZG_CUDA_TRY(cudaMemcpyAsync(d_bytes, g1_lagrange_host, 96, cudaMemcpyHostToDevice, stream));
    monomial_form_kernel<<<1, COOP_LANES, sizeof(PairingSmem), stream>>>(d_out, d_bytes, (const G2Lines*)c->g2_lines);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaMemcpyAsync(is_monomial, d_out, sizeof(int), cudaMemcpyDeviceToHost, stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(stream));
    KZG_CUDA_TRY(cudaFreeAsync(d_bytes, stream));
    KZG_CUDA_TRY(cudaFreeAsync(d_out, stream));
    L.count(1);
    return RET_OK;
}

int launch_pairing_check(Launch& L, int* d_ok, const G1* A, const G1* B, const G1* B_extra, int line_a, int line_b) {
    pairing_check_kernel<<<1, 2 * COOP_LANES, sizeof(PairingSmem2), L.stream>>>(d_ok, A, B, B_extra, (const G2Lines*)L.ctx->g2_lines, line_a, line_b);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "pairing_check");
    return RET_OK;
}

}  // namespace kzg
