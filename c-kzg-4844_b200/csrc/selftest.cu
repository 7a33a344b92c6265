// Device self-tests behind ckzg_b200_selftest_* (include/ckzg_b200.h): they run the PTX field and
// G1 primitives on operands chosen by tests/test_gpu_units.py so the -m gpu suite can compare every
// primitive with the Python oracle.  Not on any product path.
#include "engine.h"

namespace kzg {

__global__ void selftest_field_kernel(int op, uint32_t* out, const uint32_t* a, const uint32_t* b, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op <= 3 || op == 6) {
        Fp x = to_mont<FpTag>(a + 12 * i);
        Fp y = to_mont<FpTag>(b + 12 * i);
        Fp r;
        switch (op) {
            case 0: r = mul(x, y); break;
            case 1: r = add(x, y); break;
            case 2: r = sub(x, y); break;
            case 3: r = fp_inv(x); break;
            default: r = sqr(x); break;
        }
        from_mont<FpTag>(out + 12 * i, r);
    } else {
        Fr x = to_mont<FrTag>(a + 8 * i);
        Fr y = to_mont<FrTag>(b + 8 * i);
        Fr r = (op == 4) ? mul(x, y) : fr_inv(x);
        from_mont<FrTag>(out + 8 * i, r);
    }
}

__global__ void selftest_g1_kernel(int op, uint8_t* out48, int* ok_out, const uint8_t* p48, const uint32_t* k, const uint8_t* q48, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[48];
    for (int j = 0; j < 48; j++) buf[j] = p48[48 * i + j];
    G1Affine p;
    bool ok;
    if (op == 1)
        ok = g1a_validate(p, buf);
    else
        ok = g1a_uncompress(p, buf);
    G1Affine r = p;
    if (ok && op == 0) {
        uint32_t kk[8];
        for (int j = 0; j < 8; j++) kk[j] = k[8 * i + j];
        G1 acc = g1_mul_affine<8>(p, kk);
        if (q48) {
            G1Affine q;
            for (int j = 0; j < 48; j++) buf[j] = q48[48 * i + j];
            ok = g1a_uncompress(q, buf);
            if (ok) g1_madd_to(acc, q, false);
        }
        r = g1_to_affine(acc);
    }
    ok_out[i] = ok ? 1 : 0;
    if (ok) {
        g1a_compress(buf, r);
        for (int j = 0; j < 48; j++) out48[48 * i + j] = buf[j];
    }
}

// Montgomery-multiplier throughput probe: ILP independent dependent-chains per thread.
// `active` < 32: only the first `active` lanes of every warp work (does a partial warp issue faster?)
template <int ILP>
__global__ void mulbench_kernel(uint32_t* out, const uint32_t* in, int iters, int active) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(threadIdx.x & 31) >= active) return;
    Fp x[ILP], y;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < 12; i++) x[k].l[i] = in[(i + k) % 12] + tid + k;
#pragma unroll
    for (int i = 0; i < 12; i++) y.l[i] = in[12 + i];
    y.l[11] &= 0x0fffffffu;
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k].l[11] &= 0x0fffffffu;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = mul(x[k], y);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < 12; i++) acc ^= x[k].l[i];
    out[tid] = acc;
}

// FP64 FMA issue rate (8 independent chains per thread): is the DFMA pipe an alternative multiplier?
__global__ void dfmabench_kernel(double* out, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double a[8], b = 1.0000001, c = 1e-9 * tid;
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + k + tid;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[tid] = s;
}

// ilp: bits 0..7 = independent chains per thread, bits 8..15 = active lanes per warp (0 = all 32),
// bit 16 = run the DFMA probe instead (8 FMAs per thread per iteration)
int selftest_mulbench(int ilp_arg, int iters, int blocks, int threads, float* ms_out) {
    const int ilp = ilp_arg & 0xff;
    const int active = ((ilp_arg >> 8) & 0xff) ? ((ilp_arg >> 8) & 0xff) : 32;
    const bool dfma = (ilp_arg >> 16) & 1;
    uint32_t h_in[24];
    for (int i = 0; i < 24; i++) h_in[i] = 0x9e3779b9u * (i + 1);
    uint32_t *d_in = nullptr, *d_out = nullptr;
    KZG_CUDA_TRY(cudaMalloc((void**)&d_in, sizeof(h_in)));
    KZG_CUDA_TRY(cudaMalloc((void**)&d_out, (size_t)blocks * threads * 8));
    KZG_CUDA_TRY(cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (dfma) dfmabench_kernel<<<blocks, threads>>>((double*)d_out, iters);
        else if (ilp == 1) mulbench_kernel<1><<<blocks, threads>>>(d_out, d_in, iters, active);
        else if (ilp == 2) mulbench_kernel<2><<<blocks, threads>>>(d_out, d_in, iters, active);
        else mulbench_kernel<4><<<blocks, threads>>>(d_out, d_in, iters, active);
        cudaEventRecord(e1);
        KZG_CUDA_TRY(cudaEventSynchronize(e1));
    }
    KZG_CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_in);
    cudaFree(d_out);
    return RET_OK;
}

template <class T>
static int dev_copy_in(T** d, const T* h, size_t count) {
    if (!h) {
        *d = nullptr;
        return RET_OK;
    }
    KZG_CUDA_TRY(cudaMalloc((void**)d, count * sizeof(T)));
    KZG_CUDA_TRY(cudaMemcpy(*d, h, count * sizeof(T), cudaMemcpyHostToDevice));
    return RET_OK;
}

int selftest_field(int op, uint32_t* out, const uint32_t* a, const uint32_t* b, uint64_t n) {
    int limbs = (op <= 3 || op == 6) ? 12 : 8;
    uint32_t *da = nullptr, *db = nullptr, *dout = nullptr;
    int rc;
    if ((rc = dev_copy_in(&da, a, n * limbs))) return rc;
    if ((rc = dev_copy_in(&db, b, n * limbs))) return rc;
    KZG_CUDA_TRY(cudaMalloc((void**)&dout, n * limbs * sizeof(uint32_t)));
    selftest_field_kernel<<<(unsigned)((n + 63) / 64), 64>>>(op, dout, da, db, n);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaMemcpy(out, dout, n * limbs * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return RET_OK;
}

int selftest_g1(int op, uint8_t* out48, int* ok_out, const uint8_t* p48, const uint32_t* k, const uint8_t* q48, uint64_t n) {
    uint8_t *dp = nullptr, *dq = nullptr, *dout = nullptr;
    uint32_t* dk = nullptr;
    int* dok = nullptr;
    int rc;
    if ((rc = dev_copy_in(&dp, p48, n * 48))) return rc;
    if ((rc = dev_copy_in(&dq, q48, n * 48))) return rc;
    if ((rc = dev_copy_in(&dk, k, n * 8))) return rc;
    KZG_CUDA_TRY(cudaMalloc((void**)&dout, n * 48));
    KZG_CUDA_TRY(cudaMemset(dout, 0, n * 48));
    KZG_CUDA_TRY(cudaMalloc((void**)&dok, n * sizeof(int)));
    selftest_g1_kernel<<<(unsigned)((n + 31) / 32), 32>>>(op, dout, dok, dp, dk, dq, n);
    KZG_CUDA_TRY(cudaGetLastError());
    KZG_CUDA_TRY(cudaMemcpy(out48, dout, n * 48, cudaMemcpyDeviceToHost));
    KZG_CUDA_TRY(cudaMemcpy(ok_out, dok, n * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dp);
    cudaFree(dq);
    cudaFree(dk);
    cudaFree(dout);
    cudaFree(dok);
    return RET_OK;
}

}  // namespace kzg
