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
// The workspace (lane schedules, registers, all 2 x 68 evaluated lines) is dynamic shared memory.
struct PairingSmem {
    CoopWS ws;
    CoopLines ln;
    G1 pts[2];
    int dec_ok[2];
};
// pairing_check_kernel: up to four cooperative machines, one line store
struct PairingSmem4 {
    CoopWS ws[4];
    CoopLines ln;
    G1 pts[2];
};
static_assert(sizeof(PairingSmem4) <= 227 * 1024, "the four-machine workspace must fit the 227 KB of opt-in shared memory of one CTA");
extern __shared__ __align__(16) unsigned char pairing_smem_raw[];

// the lane schedules in the form the kernels keep in shared memory, built once per context (coop_init_from copies them)
__global__ void __launch_bounds__(COOP_LANES) pairing_tables_kernel(CoopTables* out) {
    CoopWS& ws = *reinterpret_cast<CoopWS*>(pairing_smem_raw);
    coop_init_tables(ws);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&ws.tb);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out);
    for (int i = threadIdx.x; i < (int)(sizeof(CoopTables) / 4); i += COOP_LANES) dst[i] = src[i];
}

__global__ void __launch_bounds__(COOP_LANES) monomial_form_kernel(int* out, const uint8_t* lag01, const G2Lines* lines, const CoopTables* tables) {
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
    coop_pairing_product_is_one(ws, &sm.ln, pts[1], &lines[0], pts[0], &lines[1], true, tables);
    if (threadIdx.x == 0) *out = (dec_ok[0] && dec_ok[1]) ? ws.result : 0;  // decoding errors are reported by the main pass
}

// ok = [ e(A, Q_a) == e(B + B_extra, Q_b) ] = [ e(-A, Q_a) * e(B + B_extra, Q_b) == 1 ];  A, B XYZZ sums.
// One pairing check is ~450 DEPENDENT tower operations of 2.4-3.3 us: the kernel is built around the length of that chain.
//  * 256 threads = FOUR machines (mode bit 3): the Miller loop of both pairings as one chain of 61 squarings and 13
//    products on machine 0, fed by three helpers that multiply the lines together group by group (coop_miller_quad);
//    128 threads = two machines, one Miller loop each (63 squarings + 68 line products), values multiplied afterwards.
//  * final exponentiation on machines 0 and 1 (mode bit 0): machine 0 squares, machine 1 multiplies
//    (coop_final_exp_is_one_duo); otherwise machine 0 alone.
// mode bits 1, 2: second form of the cyclotomic square, packed-code / tree-sum form of the other operations.
// ticks / reps: measurement hook (ckzg_b200_debug_pairing_probe), nullptr / 0 on the product path.
__global__ void __launch_bounds__(4 * COOP_LANES) pairing_check_kernel(int* ok, const G1* A, const G1* B, const G1* B_extra, const uint8_t* two48, const G2Lines* lines, int line_a,
                                                                       int line_b, int mode, long long* ticks, int reps, const CoopTables* tables) {
    PairingSmem4& sm = *reinterpret_cast<PairingSmem4*>(pairing_smem_raw);
    const int g = threadIdx.x >> 6, nm = blockDim.x >> 6;
    const bool duo = mode & 1, quad = nm == 4;
    CoopWS& ws = sm.ws[g];
    G1* pts = sm.pts;
    if (two48) {  // probe: two compressed points
        if (threadIdx.x < 2) {
            uint8_t buf[48];
            G1Affine a;
            for (int k = 0; k < 48; k++) buf[k] = two48[48 * threadIdx.x + k];
            g1a_uncompress(a, buf);
            pts[threadIdx.x] = g1_from_affine(a);
        }
    } else if (threadIdx.x == 0) {
        pts[0] = *A;
        G1 b = *B;
        if (B_extra) {
            G1 e = *B_extra;
            g1_add_to(b, e);
        }
        pts[1] = b;
    }
    __syncthreads();
    if (ticks && threadIdx.x == 0) ticks[0] = clock64();
    if (tables) coop_init_from(ws, tables);
    else coop_init_tables(ws);
    coop_attach_lines(ws, &sm.ln);
    coop_load_points(ws, pts[0], &lines[line_a], pts[1], &lines[line_b], true);
    if ((threadIdx.x & (COOP_LANES - 1)) == 0 && !quad) ws.use[1 - g] = 0;  // two machines: machine g owns pairing g
    if ((threadIdx.x & (COOP_LANES - 1)) == 1) ws.cyc2 = (mode >> 1) & 1;
    if ((threadIdx.x & (COOP_LANES - 1)) == 2) ws.run2 = (mode >> 2) & 1;
    coop_sync();
    if (quad) coop_prepare_all_lines(ws, &lines[line_a], &lines[line_b], g, 4);  // every machine a quarter of both pairs
    else coop_prepare_all_lines(ws, &lines[line_a], &lines[line_b]);             // its own pair
    __syncthreads();
    if (ticks && threadIdx.x == 0) ticks[1] = clock64();
    if (quad) coop_miller_quad(sm.ws, g);
    else {
        coop_miller_loop(ws);
        __syncthreads();
    }
    if (g >= 2 || (g == 1 && !duo)) return;
    if (g == 0) {
        const int lane = threadIdx.x;
        if (ticks && lane == 0) ticks[2] = clock64();
        if (lane == 0) ws.ticks = ticks;
        if (!quad) {  // F = F_0 * F_1
            if (lane < 12) {
                ws.reg[6][lane] = sm.ws[1].reg[0][lane];
                ws.nreg[6][lane] = sm.ws[1].nreg[0][lane];
            }
            coop_sync();
            coop_mul(ws, 0, 0, 6);
        } else {
            coop_sync();
        }
        COOP_TICK(ws, 3);
    }
    if (duo) coop_final_exp_is_one_duo(sm.ws[0], sm.ws[1], g);
    else coop_final_exp_is_one(ws);
    if (g != 0) return;
    COOP_TICK(ws, 9);
    if (threadIdx.x == 0) *ok = ws.result;
    if (!ticks || !reps) return;
    // probe: per-operation loops on register 1 (E: a cyclotomic-subgroup element after the easy part)
    for (int k = 0; k < 8; k++) {
        COOP_TICK(ws, 16 + 2 * k);
        for (int r = 0; r < reps; r++) {
            if (k == 0) coop_cyc(ws, 2, 2 - (r == 0));
            else if (k == 1) coop_mul(ws, 3, r == 0 ? 1 : 3, 1);
            else if (k == 2) coop_sqr(ws, 4, r == 0 ? 1 : 4);
            else if (k == 3) coop_line(ws, 5, r == 0 ? 1 : 5, 0, r % MILLER_LINES);
            else if (k == 4) coop_conj(ws, 6, 1);
            else if (k == 5) coop_copy(ws, 6, 1);
            else if (k == 6) coop_frobenius(ws, 6, 1, 1 + (r & 1));
            else if (r < 2) {
                if (duo) coop_inv_norm(ws, 6, 1, 7, 5);
                else coop_inv(ws, 6, 1);
            }
        }
        COOP_TICK(ws, 17 + 2 * k);
    }
}

// > 48 KB of dynamic shared memory needs the per-function opt-in, once per device context
static int pairing_smem_opt_in() {
    KZG_CUDA_TRY(cudaFuncSetAttribute(monomial_form_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairingSmem)));
    KZG_CUDA_TRY(cudaFuncSetAttribute(pairing_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairingSmem4)));
    KZG_CUDA_TRY(cudaFuncSetAttribute(pairing_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CoopWS)));
    return RET_OK;
}
// CKZG_B200_PAIRING_MODE (A/B): bit 0 = final exponentiation on two machines (0: machine 0 alone, serial inversion),
// bit 1 = second form of the cyclotomic square, bit 2 = packed term codes / tree sums in the other operations,
// bit 3 = Miller loop on four machines (0: two machines, one pairing each).  Default 15.
static int pairing_mode() {
    static const int mode = getenv("CKZG_B200_PAIRING_MODE") ? (atoi(getenv("CKZG_B200_PAIRING_MODE")) & 15) : 15;
    return mode;
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
    KZG_CUDA_TRY(cudaMalloc(&c->pairing_tables, sizeof(CoopTables)));
    pairing_tables_kernel<<<1, COOP_LANES, sizeof(CoopWS), stream>>>((CoopTables*)c->pairing_tables);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaFreeAsync(d_bytes, stream));
    L.count(3);
    return RET_OK;
}

int setup_is_monomial_form(cudaStream_t stream, Launch& L, Ctx* c, const uint8_t* g1_lagrange_host, int* is_monomial) {
    uint8_t* d_bytes = nullptr;
    int* d_out = nullptr;
    KZG_CUDA_TRY(cudaMallocAsync((void**)&d_bytes, 96, stream));
    KZG_CUDA_TRY(cudaMallocAsync((void**)&d_out, sizeof(int), stream));
    KZG_CUDA_TRY(cudaMemcpyAsync(d_bytes, g1_lagrange_host, 96, cudaMemcpyHostToDevice, stream));
    monomial_form_kernel<<<1, COOP_LANES, sizeof(PairingSmem), stream>>>(d_out, d_bytes, (const G2Lines*)c->g2_lines, (const CoopTables*)c->pairing_tables);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaMemcpyAsync(is_monomial, d_out, sizeof(int), cudaMemcpyDeviceToHost, stream));
    KZG_CUDA_TRY(cudaStreamSynchronize(stream));
    KZG_CUDA_TRY(cudaFreeAsync(d_bytes, stream));
    KZG_CUDA_TRY(cudaFreeAsync(d_out, stream));
    L.count(1);
    return RET_OK;
}

int launch_pairing_check(Launch& L, int* d_ok, const G1* A, const G1* B, const G1* B_extra, int line_a, int line_b) {
    const int mode = pairing_mode();
    pairing_check_kernel<<<1, ((mode & 8) ? 4 : 2) * COOP_LANES, sizeof(PairingSmem4), L.stream>>>(d_ok, A, B, B_extra, nullptr, (const G2Lines*)L.ctx->g2_lines, line_a, line_b, mode,
                                                                                                 nullptr, 0, (const CoopTables*)L.ctx->pairing_tables);
    KZG_CUDA_TRY(cudaGetLastError());
    L.count(1, "pairing_check");
    return RET_OK;
}

// ticks_host: 64 values (see pairing_check_kernel); two48_host: two compressed G1 points
int debug_pairing_probe(Ctx* c, long long* ticks_host, int* ok_host, const uint8_t* two48_host, int reps) {
    const int mode = getenv("CKZG_B200_PAIRING_MODE") ? (atoi(getenv("CKZG_B200_PAIRING_MODE")) & 15) : 15;
    int prev = 0;
    cudaGetDevice(&prev);
    KZG_CUDA_TRY(cudaSetDevice(c->device));
    long long* d_t = nullptr;
    int* d_ok = nullptr;
    uint8_t* d_p = nullptr;
    KZG_CUDA_TRY(cudaMalloc((void**)&d_t, 64 * sizeof(long long)));
    KZG_CUDA_TRY(cudaMemset(d_t, 0, 64 * sizeof(long long)));
    KZG_CUDA_TRY(cudaMalloc((void**)&d_ok, sizeof(int)));
    KZG_CUDA_TRY(cudaMalloc((void**)&d_p, 96));
    KZG_CUDA_TRY(cudaMemcpy(d_p, two48_host, 96, cudaMemcpyHostToDevice));
    pairing_check_kernel<<<1, ((mode & 8) ? 4 : 2) * COOP_LANES, sizeof(PairingSmem4)>>>(d_ok, nullptr, nullptr, nullptr, d_p, (const G2Lines*)c->g2_lines, 0, 1, mode, d_t, reps,
                                                                                         getenv("CKZG_B200_PAIRING_TABLES_INLINE") ? nullptr : (const CoopTables*)c->pairing_tables);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaMemcpy(ticks_host, d_t, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
    KZG_CUDA_TRY(cudaMemcpy(ok_host, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_t);
    cudaFree(d_ok);
    cudaFree(d_p);
    cudaSetDevice(prev);
    return RET_OK;
}

}  // namespace kzg
