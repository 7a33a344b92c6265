/*
 * ckzg.h -- the frozen c-kzg-4844 C API, served by the B200 engine.
 *
 * Drop-in for the reference's `src/ckzg.h` umbrella (which pulls in src/eip4844/eip4844.h:43-84,
 * src/eip7594/eip7594.h:35-68, src/setup/setup.h:31-44, src/common/bytes.h, src/common/ret.h:24-29,
 * src/eip4844/blob.h, src/eip7594/cell.h, src/setup/settings.h:27-79): same type names, same
 * function names, same argument order and meaning, same C_KZG_RET codes, same 80-byte
 * caller-allocated KZGSettings.  A binding that today compiles the reference's ckzg.c instead links
 * libckzg_b200.so and includes this header (INTEGRATION.md shows the cgo/JNI/N-API/ctypes stubs).
 *
 * The forwarding headers include/eip4844/eip4844.h, include/eip7594/eip7594.h, include/setup/setup.h
 * and include/common/{bytes,ret}.h exist so that `#include "eip4844/eip4844.h"`-style includes keep
 * working.  blst is not needed: fr_t / g1_t / g2_t are layout-compatible opaque structs here.
 */
#ifndef CKZG_H
#define CKZG_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (src/eip4844/blob.h:29-42, src/eip7594/cell.h:28-37, src/common/bytes.h:29-42) ---- */
#define BYTES_PER_COMMITMENT 48
#define BYTES_PER_PROOF 48
#define BYTES_PER_FIELD_ELEMENT 32
#define BITS_PER_FIELD_ELEMENT 255
#define FIELD_ELEMENTS_PER_BLOB 4096
#define BYTES_PER_BLOB (FIELD_ELEMENTS_PER_BLOB * BYTES_PER_FIELD_ELEMENT)
#define LOG_EXPANSION_FACTOR 1
#define FIELD_ELEMENTS_PER_EXT_BLOB (FIELD_ELEMENTS_PER_BLOB << LOG_EXPANSION_FACTOR)
#define FIELD_ELEMENTS_PER_CELL 64
#define BYTES_PER_CELL (FIELD_ELEMENTS_PER_CELL * BYTES_PER_FIELD_ELEMENT)
#define CELLS_PER_BLOB (FIELD_ELEMENTS_PER_BLOB / FIELD_ELEMENTS_PER_CELL)
#define CELLS_PER_EXT_BLOB (FIELD_ELEMENTS_PER_EXT_BLOB / FIELD_ELEMENTS_PER_CELL)
#define BYTES_PER_G1 48
#define BYTES_PER_G2 96
#define NUM_G1_POINTS FIELD_ELEMENTS_PER_BLOB
#define NUM_G2_POINTS 65

/* ---- return codes (src/common/ret.h:24-29) ---- */
typedef enum {
    C_KZG_OK = 0,  /* Success */
    C_KZG_BADARGS, /* The supplied data is invalid in some way */
    C_KZG_ERROR,   /* Internal error (here also: no usable CUDA device -- there is no CPU path) */
    C_KZG_MALLOC,  /* Could not allocate (host or device) memory */
} C_KZG_RET;

/* ---- wire types ---- */
typedef struct { uint8_t bytes[32]; } Bytes32;            /* src/common/bytes.h:49 */
typedef struct { uint8_t bytes[48]; } Bytes48;            /* src/common/bytes.h:54 */
typedef struct { uint8_t bytes[BYTES_PER_BLOB]; } Blob;   /* src/eip4844/blob.h:49 */
typedef struct { uint8_t bytes[BYTES_PER_CELL]; } Cell;   /* src/eip7594/cell.h:44 */
typedef Bytes48 KZGCommitment;                            /* src/eip4844/eip4844.h:30 */
typedef Bytes48 KZGProof;                                 /* src/eip4844/eip4844.h:33 */

/* Same sizes as blst_fr / blst_p1 / blst_p2 (32 / 144 / 288 bytes).  Contents are private to this
 * library: an fr_t holds the canonical scalar, a g1_t the validated compressed point (see ckzg.c). */
typedef struct { uint64_t l[4]; } fr_t;
typedef struct { uint64_t l[18]; } g1_t;
typedef struct { uint64_t l[36]; } g2_t;

/*
 * KZGSettings (src/setup/settings.h:27-79): caller-allocated, 8 pointers + 2 size_t = 80 bytes.
 * The setup lives in GPU memory; the reference's host arrays are therefore not populated.  The
 * pointer slots keep their names; `tables` carries the engine context handle.
 */
typedef struct {
    fr_t *roots_of_unity;          /* NULL: device resident */
    fr_t *brp_roots_of_unity;      /* NULL: device resident */
    fr_t *reverse_roots_of_unity;  /* NULL: device resident */
    g1_t *g1_values_monomial;      /* NULL: device resident */
    g1_t *g1_values_lagrange_brp;  /* NULL: device resident */
    g2_t *g2_values_monomial;      /* NULL: device resident */
    g1_t **x_ext_fft_columns;      /* NULL: device resident */
    void **tables;                 /* engine context (ckzg_b200_ctx *), owned by the library */
    size_t wbits;                  /* the `precompute` argument, as in the reference */
    size_t scratch_size;           /* 0 */
} KZGSettings;

/* ---- setup (src/setup/setup.h:31-44) ---- */
C_KZG_RET load_trusted_setup(
    KZGSettings *out,
    const uint8_t *g1_monomial_bytes,
    uint64_t num_g1_monomial_bytes,
    const uint8_t *g1_lagrange_bytes,
    uint64_t num_g1_lagrange_bytes,
    const uint8_t *g2_monomial_bytes,
    uint64_t num_g2_monomial_bytes,
    uint64_t precompute
);
C_KZG_RET load_trusted_setup_file(KZGSettings *out, FILE *in, uint64_t precompute);
void free_trusted_setup(KZGSettings *s);

/* ---- EIP-4844 (src/eip4844/eip4844.h:43-81) ---- */
C_KZG_RET blob_to_kzg_commitment(KZGCommitment *out, const Blob *blob, const KZGSettings *s);
C_KZG_RET compute_kzg_proof(
    KZGProof *proof_out, Bytes32 *y_out, const Blob *blob, const Bytes32 *z_bytes, const KZGSettings *s
);
C_KZG_RET compute_blob_kzg_proof(
    KZGProof *out, const Blob *blob, const Bytes48 *commitment_bytes, const KZGSettings *s
);
C_KZG_RET verify_kzg_proof(
    bool *ok,
    const Bytes48 *commitment_bytes,
    const Bytes32 *z_bytes,
    const Bytes32 *y_bytes,
    const Bytes48 *proof_bytes,
    const KZGSettings *s
);
C_KZG_RET verify_blob_kzg_proof(
    bool *ok, const Blob *blob, const Bytes48 *commitment_bytes, const Bytes48 *proof_bytes, const KZGSettings *s
);
C_KZG_RET verify_blob_kzg_proof_batch(
    bool *ok,
    const Blob *blobs,
    const Bytes48 *commitments_bytes,
    const Bytes48 *proofs_bytes,
    uint64_t n,
    const KZGSettings *s
);

/* ---- EIP-7594 (src/eip7594/eip7594.h:35-56) ---- */
C_KZG_RET compute_cells_and_kzg_proofs(Cell *cells, KZGProof *proofs, const Blob *blob, const KZGSettings *s);
C_KZG_RET recover_cells_and_kzg_proofs(
    Cell *recovered_cells,
    KZGProof *recovered_proofs,
    const uint64_t *cell_indices,
    const Cell *cells,
    uint64_t num_cells,
    const KZGSettings *s
);
C_KZG_RET verify_cell_kzg_proof_batch(
    bool *ok,
    const Bytes48 *commitments_bytes,
    const uint64_t *cell_indices,
    const Cell *cells,
    const Bytes48 *proofs_bytes,
    uint64_t num_cells,
    const KZGSettings *s
);

/* ---- helpers the reference exports and its Go/Rust tests call (src/common/bytes.h:66-74,
 *      src/eip4844/eip4844.h:84, src/eip7594/eip7594.h:58-68) ---- */
void bytes_from_uint64(uint8_t out[8], uint64_t n);
void bytes_from_g1(Bytes48 *out, const g1_t *in);
void bytes_from_bls_field(Bytes32 *out, const fr_t *in);
C_KZG_RET bytes_to_bls_field(fr_t *out, const Bytes32 *b);
C_KZG_RET bytes_to_kzg_commitment(g1_t *out, const Bytes48 *b);
C_KZG_RET bytes_to_kzg_proof(g1_t *out, const Bytes48 *b);
void compute_challenge(fr_t *eval_challenge_out, const Blob *blob, const g1_t *commitment);
C_KZG_RET compute_verify_cell_kzg_proof_batch_challenge(
    fr_t *challenge_out,
    const Bytes48 *commitments_bytes,
    uint64_t num_commitments,
    const uint64_t *commitment_indices,
    const uint64_t *cell_indices,
    const Cell *cells,
    const Bytes48 *proofs_bytes,
    uint64_t num_cells
);

#ifdef __cplusplus
}
#endif
#endif /* CKZG_H */
