// Internal interface of the EIP-7594 side of the engine (cells.cu, fk20.cu, recover.cu).
#pragma once
#include "engine.h"

namespace kzg {

constexpr int CELLS_EXT = 128;        // CELLS_PER_EXT_BLOB (src/eip7594/cell.h:37)
constexpr int CELL_FR = 64;           // FIELD_ELEMENTS_PER_CELL (cell.h:28)
constexpr int CELL_BYTES = 2048;      // BYTES_PER_CELL
constexpr int FK_POINTS = CELLS_EXT * CELL_FR;  // 8192 fixed bases X^[j][i] (x_ext_fft_columns, setup.c:272-289)

// Fixed-base tables for the 128 x MSM(64) of FK20: T[p][w][m] = (m+1) * 2^(c w) * X^_p, affine,
// p = j*64 + i, w < ceil(256 / c) windows of c signed bits, m < 2^(c-1).  HBM buys a pure gather-and-add
// MSM: 64 * ceil(256 / c) mixed additions, no buckets, no doublings (the role the reference gives to
// `precompute`, setup.c:291-323 / README.md:110-143 -- 96 MiB at precompute=8).  The window width is
// chosen from the device memory that is free when the table is first needed (api.cu plan_fk_window): c = 12 -> 22 windows,
// 35 GB (1408 additions per MSM); c = 10 -> 26 windows, 10.5 GB; c = 8 -> 32 windows, 3.2 GB (2048).
struct FkGeom {
    int c = 8, w = 32, m = 128;
    size_t table_points() const { return (size_t)FK_POINTS * w * m; }
    size_t points_for(size_t npts) const { return npts * (size_t)w * m; }
};
inline FkGeom fk_geom(int c) {
    FkGeom g;
    g.c = c;
    g.w = (256 + c - 1) / c;
    g.m = 1 << (c - 1);
    return g;
}

// ---- cells.cu ------------------------------------------------------------------------------------
// cells (n x 128 x 2048 B, may be null) and/or monomial coefficients (n x 4096 Fr, may be null)
int launch_blob_to_cells(Launch& L, uint8_t* cells, Fr* mono, const uint8_t* blobs, uint64_t n, int* d_bad);
// S[blob][j][i] = plain limbs of FFT128(c_i)[j] / 128  (fk20.c:199-209)
int launch_fk20_scalars(Launch& L, uint32_t* S, const Fr* mono, uint64_t n);

// ---- fk20.cu -------------------------------------------------------------------------------------
// X^ columns (init_fk20_multi_settings, setup.c:238-330) + the window tables, built on first use
int fk20_ensure(Ctx* c);
// table[(p*W + w)*M + m] = (m+1) 2^(cw) pts[p], affine, for npts fixed bases
int launch_fixed_base_table(Launch& L, G1Affine* table, const G1Affine* pts, int npts, const FkGeom& g);
// u_brp[blob][brp7(j)] = sum_i S[blob][j][i] * X^[j][i]
int launch_fk20_msm(Launch& L, G1* u_brp, const uint32_t* S, uint64_t n);
// the same sums by pairwise affine additions with batched inversions (msm_affine.cu), for batches
size_t fk20_msm_affine_workspace_bytes(uint64_t n, int c);
int launch_fk20_msm_affine(Launch& L, G1* u_brp, const uint32_t* S, uint64_t n, void* workspace);
// proofs[blob][128] (XYZZ, final bit-reversed order) from u_brp: unscaled inverse G1 FFT, zero the
// upper half, forward G1 FFT (fk20.c:257-269 + eip7594.c:133)
int launch_fk20_g1_ffts(Launch& L, G1* proofs, G1* u_brp, uint64_t n);

// ---- fk20_fft.cu ---------------------------------------------------------------------------------
// in place, nvec vectors of 128 XYZZ points: [unscaled inverse FFT of bit-reversed input, lower half kept ->]
// forward FFT with the upper half taken as infinity, output bit-reversed
int g1_fft128_run(Launch& L, G1* data, uint64_t nvec, bool with_inverse);

// ---- recover.cu ----------------------------------------------------------------------------------
int recover_setup(Launch& L, Ctx* c);
size_t recover_scratch_bytes(uint64_t n);
// recover_cells (recovery.c:200): cells_out n x 128 x 2048 B (device), mono_out n x 4096 Fr or null
int launch_recover(Launch& L, uint8_t* cells_out, Fr* mono_out, const uint8_t* cells_in, const int16_t* slot, const uint8_t* present, uint64_t num_cells, uint64_t n, int* d_bad,
                   void* scratch);

// ---- verify_cells.cu -----------------------------------------------------------------------------
// Table columns of one call: proofs 0..n-1, unique commitments n..n+u-1, then the 64 (negated) setup
// points of the interpolation commitment; VMSM_LEVELS rows (vmsm.cu layout).
int setup_verify_cells(Launch& L, Ctx* c);  // Ctx::mono_levels
size_t verify_cells_table_points(uint64_t n, uint64_t u);
// before the challenge: decompress + subgroup-check proofs and unique commitments (bytes_to_kzg_proof
// eip7594.c:917-920, bytes_to_kzg_commitment :513), leaving their table columns; copies the fixed columns
int launch_verify_cells_validate(Launch& L, G1* table, const uint8_t* proofs48, uint64_t n, const uint8_t* uniq48, uint64_t u, int* d_bad);
size_t verify_cells_scratch_bytes(uint64_t n, uint64_t u);
// out2[0] = sum r^k pi_k, out2[1] = sum w_c C_c - [I] + sum r^k h_k^64 pi_k  (eip7594.c:825-974);
// r = hash_to_bls_field(digest32) (HOST pointer); col_of[k] = cell index of cell k (device, n bytes)
int launch_verify_cells(Launch& L, G1* out2, const G1* table, const uint8_t* cells, const uint8_t* digest32, const uint8_t* col_of, const uint32_t* col_start,
                        const uint32_t* col_items, const uint32_t* cm_start, const uint32_t* cm_items, uint64_t n, uint64_t u, int* d_bad, void* scratch);

}  // namespace kzg
