// Internal interface of the proof / verification side of the engine (pairing.cu, verify.cu).
#pragma once
#include "engine.h"

namespace kzg {

// line-table slots in Ctx::g2_lines
enum { LINE_G2_GEN = 0, LINE_G2_TAU = 1, LINE_G2_TAU64 = 2 };

// ---- pairing.cu ----------------------------------------------------------------------------------
// Decompress the 65 G2 points (setup.c:469-477; failure sets *d_bad) and precompute the Miller-loop
// lines of the fixed G2 arguments (G2 generator, [tau]G2, [tau^64]G2).
int setup_g2_and_lines(cudaStream_t stream, Launch& L, Ctx* c, const uint8_t* g2_monomial_bytes_host, int* d_bad);
// is_trusted_setup_in_lagrange_form (setup.c:339-358): pairing check on the first two Lagrange points
// (in file order, i.e. before the bit-reversal).
int setup_is_monomial_form(cudaStream_t stream, Launch& L, Ctx* c, const uint8_t* g1_lagrange_bytes_host, int* is_monomial);
// *d_ok = [ e(A, Q_line_a) == e(B + B_extra, Q_line_b) ]
int launch_pairing_check(Launch& L, int* d_ok, const G1* A, const G1* B, const G1* B_extra, int line_a, int line_b);

// ---- verify.cu -----------------------------------------------------------------------------------
// z[i] = hash_to_bls_field(SHA256("FSBLOBVERIFY_V1_" || 0 || 4096 || blob_i || commitment_i))
// (compute_challenge, src/eip4844/eip4844.c:147).  zy[i*64 .. +32) receives canonical z bytes.
int launch_blob_challenges(Launch& L, Fr* z, uint8_t* zy, const uint8_t* blobs, const uint8_t* commitments48, uint64_t n);
// the same hash in pieces: SHA-256 blocks [k0, k1) of blob_challenge_blocks() = 2050 (block k >= 1 covers blob bytes
// [64k - 32, 64k + 32)), chaining values carried in states (8 words per blob)
int blob_challenge_blocks();
int launch_blob_challenges_range(Launch& L, Fr* z, uint8_t* zy, const uint8_t* blobs, const uint8_t* commitments48, uint64_t n, int k0, int k1, uint32_t* states);
// z[i] from canonical bytes (compute_kzg_proof path); bad[i] = 1 if >= r
int launch_z_from_bytes(Launch& L, Fr* z, uint8_t* zy, const uint8_t* z_bytes, uint64_t n, int* bad);
// y[i] = p_i(z[i]) (evaluate_polynomial_in_evaluation_form, eip4844.c:192), canonical y bytes into
// zy[i*64+32 ..).  bad[i*bad_stride] = 1 for a non-canonical blob element.  Optionally stores the
// 4096 inverses 1/(z - w_j) per blob and the in-domain index (or -1).
int launch_evaluate(Launch& L, Fr* y, uint8_t* zy, Fr* inv_or_null, int* m_or_null, const uint8_t* blobs, const Fr* z, uint64_t n, int* bad, int bad_stride);
// quotient scalars (plain little-endian limbs) for compute_kzg_proof_impl (eip4844.c:417-494)
int launch_quotient(Launch& L, uint8_t* q_scalars, const uint8_t* blobs, const Fr* z, const Fr* y, const Fr* inv, const int* m, uint64_t n);
// decompress + validate n points (validate_kzg_g1, bytes.c:81); bad[i*bad_stride] = 1 on failure
int launch_g1_validate(Launch& L, G1Affine* out, const uint8_t* bytes48, uint64_t n, int* bad, int bad_stride);
// same for two arrays of n points in one launch (single shared flag)
int launch_g1_validate2(Launch& L, G1Affine* out_a, const uint8_t* in_a, G1Affine* out_b, const uint8_t* in_b, uint64_t n, int* bad);
// hash (z) of n blobs and validation of their n commitments + n proofs in one launch (4 warps per 32 blobs)
// table != nullptr: also leaves the vmsm table columns of the 2n points (layout above)
int launch_stage1_fused(Launch& L, Fr* z, uint8_t* zy, const uint8_t* blobs, G1Affine* out_cm, const uint8_t* in_cm, G1Affine* out_pf, const uint8_t* in_pf, uint64_t n, int* bad,
                        G1* table);
// validation of n commitments + n proofs with the vmsm table columns
int launch_g1_validate2_levels(Launch& L, G1Affine* out_cm, const uint8_t* in_cm, G1Affine* out_pf, const uint8_t* in_pf, uint64_t n, int* bad, G1* table);
// n_a + n_b points on quads of lanes (g1_quad.cuh); table columns col_a + k / col_b + k of a table of npts columns;
// out_a / out_b (affine points) may be null
int launch_g1_validate_levels_ab(Launch& L, G1Affine* out_a, const uint8_t* in_a, uint64_t na, uint64_t col_a, G1Affine* out_b, const uint8_t* in_b, uint64_t nb, uint64_t col_b,
                                 G1* table, uint64_t npts, int* bad);
int debug_set_placement_buffer(uint32_t* dev_buf);
int debug_set_timer_buffer(uint32_t* dev_buf);
int launch_g1_validate_ab(Launch& L, G1Affine* out_a, const uint8_t* in_a, uint64_t n, G1Affine* out_b, const uint8_t* in_b, uint64_t nb, int* bad);
// r = hash_to_bls_field(digest): the batch transcript itself (eip4844.c:597-680) is hashed on the host
int launch_r_from_digest(Launch& L, Fr* r, const uint8_t* digest32);
// Random linear combination over tuples [first, first+n_local) with powers r^(first+i):
//   A = sum r^i proof_i,  B = sum r^i z_i proof_i + sum r^i C_i - [sum r^i y_i] G1
// (verify_kzg_proof_batch, eip4844.c:697-765).  use_r = false: all weights 1 (the n == 1 equation).
// Results (XYZZ, device): out[0] = A, out[1] = B.  scratch sized by rlc_scratch_bytes(n_local).
size_t rlc_scratch_bytes(uint64_t n_local);
int launch_rlc(Launch& L, G1* out2, const G1Affine* commitments, const G1Affine* proofs, const Fr* z, const Fr* y, const Fr* r, bool use_r,
               uint64_t first, uint64_t n_local, void* scratch);
// ---- vmsm.cu: the same linear combination as ONE pair of bucket MSMs over pre-shifted bases --------
// Point layout: column i < n = proof i, column n + i = commitment i, column 2n = -G1 generator.
// table[j][col], j < 9: 2^(8j) P; j = 9..17: 2^(8(j-9)) [|z|]P (XYZZ) -- written by the validation
// kernels themselves (g1.cuh g1a_validate_levels: the doubling chains of the subgroup test), i.e. before
// the challenge r exists; the generator column is copied from the context.
constexpr int VMSM_LEVELS = 18;
size_t vmsm_table_points(uint64_t n);   // VMSM_LEVELS * (2n + 1)
int launch_vmsm_generator_levels(Launch& L, G1* levels18);
int vmsm_place_generator(Launch& L, G1* table, uint64_t n);
size_t rlc_vmsm_scratch_bytes(uint64_t n);
// out2[0] = A, out2[1] = B as launch_rlc (use_r = true, weights r^(first + i)); r = hash_to_bls_field(digest32), digest32 a HOST pointer
int launch_rlc_vmsm(Launch& L, G1* out2, const G1* table, const Fr* z, const Fr* y, const uint8_t* digest32, uint64_t first, uint64_t n, void* scratch);
// sum of n XYZZ points -> out (device); in is clobbered
int launch_g1_sum(Launch& L, G1* out, G1* in, uint64_t n);
// canonical 32-byte big-endian encodings of n field elements
int launch_fr_to_bytes(Launch& L, uint8_t* out32, const Fr* in, uint64_t n);
// out[i] = XYZZ lift of in[i]
int launch_lift_affine(Launch& L, G1* out, const G1Affine* in, uint64_t n);
// Fr from 32-byte canonical big-endian (bad[i]=1 if >= r)
int launch_fr_from_bytes(Launch& L, Fr* out, const uint8_t* bytes32, uint64_t stride, uint64_t n, int* bad);

}  // namespace kzg
