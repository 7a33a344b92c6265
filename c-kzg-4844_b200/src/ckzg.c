/*
 * ckzg.c -- host C layer: the frozen c-kzg-4844 API (include/ckzg.h) on top of the engine's C ABI
 * (include/ckzg_b200.h).  Mirrors the reference's entry points one for one:
 *
 *   load_trusted_setup / _file / free_trusted_setup ... src/setup/setup.c:392,519,162
 *   blob_to_kzg_commitment ........................... src/eip4844/eip4844.c:264
 *   compute_kzg_proof / compute_blob_kzg_proof ....... src/eip4844/eip4844.c:382,506
 *   verify_kzg_proof / verify_blob_kzg_proof(_batch) . src/eip4844/eip4844.c:302,546,775
 *   compute_cells_and_kzg_proofs, recover_..., verify_cell_kzg_proof_batch
 *                                                      src/eip7594/eip7594.c:61,177,825
 *
 * This file only validates arguments the way the reference does, forwards host pointers to the
 * engine and maps return codes.  It performs no field or curve arithmetic: there is no CPU path.
 */
#include "../../include/ckzg.h"
#include "../../include/ckzg_b200.h"

#include "host_sha256.h"

#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static ckzg_b200_ctx *engine_of(const KZGSettings *s) {
    return (s == NULL) ? NULL : (ckzg_b200_ctx *)(void *)s->tables;
}

static void init_settings(KZGSettings *out) { memset(out, 0, sizeof(*out)); }

/* ---------------------------------------------------------------------------------------------- */
/* setup                                                                                          */
/* ---------------------------------------------------------------------------------------------- */

void free_trusted_setup(KZGSettings *s) {
    /* src/setup/setup.c:162-190: safe on NULL, on a zeroed struct and after a failed load */
    if (s == NULL) return;
    if (s->tables != NULL) ckzg_b200_ctx_destroy((ckzg_b200_ctx *)(void *)s->tables);
    init_settings(s);
}

C_KZG_RET load_trusted_setup(
    KZGSettings *out,
    const uint8_t *g1_monomial_bytes,
    uint64_t num_g1_monomial_bytes,
    const uint8_t *g1_lagrange_bytes,
    uint64_t num_g1_lagrange_bytes,
    const uint8_t *g2_monomial_bytes,
    uint64_t num_g2_monomial_bytes,
    uint64_t precompute
) {
    init_settings(out);
    if (precompute > 15) return C_KZG_BADARGS; /* setup.c:411 */
    if (num_g1_monomial_bytes != NUM_G1_POINTS * BYTES_PER_G1 ||
        num_g1_lagrange_bytes != NUM_G1_POINTS * BYTES_PER_G1 ||
        num_g2_monomial_bytes != NUM_G2_POINTS * BYTES_PER_G2) {
        return C_KZG_BADARGS; /* setup.c:425-430 */
    }
    ckzg_b200_ctx *ctx = NULL;
    int rc = ckzg_b200_ctx_create(&ctx, g1_monomial_bytes, g1_lagrange_bytes, g2_monomial_bytes, precompute, -1);
    if (rc != 0) {
        init_settings(out); /* setup.c:497-504: leave a state free_trusted_setup accepts */
        return (C_KZG_RET)rc;
    }
    out->tables = (void **)(void *)ctx;
    out->wbits = (size_t)precompute;
    return C_KZG_OK;
}

static int read_hex_bytes(FILE *in, uint8_t *dst, size_t n) {
    /* setup.c:558-582 reads with fscanf("%2hhx"): leading whitespace skipped, two hex digits per byte */
    for (size_t i = 0; i < n; i++) {
        if (fscanf(in, "%2hhx", &dst[i]) != 1) return 0;
    }
    return 1;
}

C_KZG_RET load_trusted_setup_file(KZGSettings *out, FILE *in, uint64_t precompute) {
    C_KZG_RET ret = C_KZG_BADARGS;
    uint64_t n1 = 0, n2 = 0;
    uint8_t *mono = NULL, *lag = NULL, *g2 = NULL;

    init_settings(out);
    mono = calloc(NUM_G1_POINTS, BYTES_PER_G1);
    lag = calloc(NUM_G1_POINTS, BYTES_PER_G1);
    g2 = calloc(NUM_G2_POINTS, BYTES_PER_G2);
    if (!mono || !lag || !g2) {
        ret = C_KZG_MALLOC;
        goto out;
    }
    if (fscanf(in, "%" SCNu64, &n1) != 1 || n1 != NUM_G1_POINTS) goto out;
    if (fscanf(in, "%" SCNu64, &n2) != 1 || n2 != NUM_G2_POINTS) goto out;
    /* file order: G1 Lagrange, G2 monomial, G1 monomial (setup.c:558-582) */
    if (!read_hex_bytes(in, lag, (size_t)NUM_G1_POINTS * BYTES_PER_G1)) goto out;
    if (!read_hex_bytes(in, g2, (size_t)NUM_G2_POINTS * BYTES_PER_G2)) goto out;
    if (!read_hex_bytes(in, mono, (size_t)NUM_G1_POINTS * BYTES_PER_G1)) goto out;
    ret = load_trusted_setup(
        out, mono, NUM_G1_POINTS * BYTES_PER_G1, lag, NUM_G1_POINTS * BYTES_PER_G1, g2, NUM_G2_POINTS * BYTES_PER_G2, precompute
    );
out:
    free(mono);
    free(lag);
    free(g2);
    return ret;
}

/* ---------------------------------------------------------------------------------------------- */
/* EIP-4844                                                                                       */
/* ---------------------------------------------------------------------------------------------- */

C_KZG_RET blob_to_kzg_commitment(KZGCommitment *out, const Blob *blob, const KZGSettings *s) {
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    return (C_KZG_RET)ckzg_b200_blob_to_kzg_commitment_coalesced(e, out->bytes, blob->bytes);
}

C_KZG_RET compute_kzg_proof(
    KZGProof *proof_out, Bytes32 *y_out, const Blob *blob, const Bytes32 *z_bytes, const KZGSettings *s
) {
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    return (C_KZG_RET)ckzg_b200_compute_kzg_proof_coalesced(e, proof_out->bytes, y_out->bytes, blob->bytes, z_bytes->bytes);
}

C_KZG_RET compute_blob_kzg_proof(
    KZGProof *out, const Blob *blob, const Bytes48 *commitment_bytes, const KZGSettings *s
) {
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    return (C_KZG_RET)ckzg_b200_compute_blob_kzg_proof_coalesced(e, out->bytes, blob->bytes, commitment_bytes->bytes);
}

C_KZG_RET verify_kzg_proof(
    bool *ok,
    const Bytes48 *commitment_bytes,
    const Bytes32 *z_bytes,
    const Bytes32 *y_bytes,
    const Bytes48 *proof_bytes,
    const KZGSettings *s
) {
    *ok = false; /* eip4844.c:314 */
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    int good = 0;
    int rc = ckzg_b200_verify_kzg_proof(e, &good, commitment_bytes->bytes, z_bytes->bytes, y_bytes->bytes, proof_bytes->bytes);
    if (rc == 0) *ok = good != 0;
    return (C_KZG_RET)rc;
}

C_KZG_RET verify_blob_kzg_proof_batch(
    bool *ok,
    const Blob *blobs,
    const Bytes48 *commitments_bytes,
    const Bytes48 *proofs_bytes,
    uint64_t n,
    const KZGSettings *s
) {
    /* eip4844.c:791-794: zero blobs verify trivially (before touching anything else) */
    if (n == 0) {
        *ok = true;
        return C_KZG_OK;
    }
    *ok = false;
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    int good = 0;
    int rc = ckzg_b200_verify_blob_kzg_proof_batch(
        e, &good, (const uint8_t *)blobs, (const uint8_t *)commitments_bytes, (const uint8_t *)proofs_bytes, n, CKZG_B200_HOST
    );
    if (rc == 0) *ok = good != 0;
    return (C_KZG_RET)rc;
}

C_KZG_RET verify_blob_kzg_proof(
    bool *ok, const Blob *blob, const Bytes48 *commitment_bytes, const Bytes48 *proof_bytes, const KZGSettings *s
) {
    *ok = false; /* eip4844.c:558 */
    return verify_blob_kzg_proof_batch(ok, blob, commitment_bytes, proof_bytes, 1, s);
}

/* ---------------------------------------------------------------------------------------------- */
/* EIP-7594                                                                                       */
/* ---------------------------------------------------------------------------------------------- */

C_KZG_RET compute_cells_and_kzg_proofs(Cell *cells, KZGProof *proofs, const Blob *blob, const KZGSettings *s) {
    if (cells == NULL && proofs == NULL) return C_KZG_BADARGS; /* eip7594.c:72-74 */
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    return (C_KZG_RET)ckzg_b200_compute_cells_and_kzg_proofs_coalesced(e, (uint8_t *)cells, (uint8_t *)proofs, blob->bytes);
}

C_KZG_RET recover_cells_and_kzg_proofs(
    Cell *recovered_cells,
    KZGProof *recovered_proofs,
    const uint64_t *cell_indices,
    const Cell *cells,
    uint64_t num_cells,
    const KZGSettings *s
) {
    /* eip7594.c:191-213: count and index checks come first */
    if (num_cells > CELLS_PER_EXT_BLOB || num_cells < CELLS_PER_BLOB) return C_KZG_BADARGS;
    for (uint64_t i = 0; i < num_cells; i++) {
        if (cell_indices[i] >= CELLS_PER_EXT_BLOB) return C_KZG_BADARGS;
        if (i > 0 && cell_indices[i] <= cell_indices[i - 1]) return C_KZG_BADARGS; /* strictly ascending */
    }
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    return (C_KZG_RET)ckzg_b200_recover_cells_and_kzg_proofs_coalesced(
        e, (uint8_t *)recovered_cells, (uint8_t *)recovered_proofs, cell_indices, (const uint8_t *)cells, num_cells
    );
}

C_KZG_RET verify_cell_kzg_proof_batch(
    bool *ok,
    const Bytes48 *commitments_bytes,
    const uint64_t *cell_indices,
    const Cell *cells,
    const Bytes48 *proofs_bytes,
    uint64_t num_cells,
    const KZGSettings *s
) {
    *ok = false; /* eip7594.c:849 */
    if (num_cells == 0) { /* eip7594.c:852-855 */
        *ok = true;
        return C_KZG_OK;
    }
    for (uint64_t i = 0; i < num_cells; i++) { /* eip7594.c:861-864 */
        if (cell_indices[i] >= CELLS_PER_EXT_BLOB) return C_KZG_BADARGS;
    }
    ckzg_b200_ctx *e = engine_of(s);
    if (e == NULL) return C_KZG_BADARGS;
    int good = 0;
    int rc = ckzg_b200_verify_cell_kzg_proof_batch(
        e, &good, (const uint8_t *)commitments_bytes, cell_indices, (const uint8_t *)cells, (const uint8_t *)proofs_bytes,
        num_cells, CKZG_B200_HOST
    );
    if (rc == 0) *ok = good != 0;
    return (C_KZG_RET)rc;
}

/* ---------------------------------------------------------------------------------------------- */
/* small helpers the reference exports (src/common/bytes.c)                                       */
/* ---------------------------------------------------------------------------------------------- */

void bytes_from_uint64(uint8_t out[8], uint64_t n) { /* bytes.c:29: big-endian */
    for (int i = 7; i >= 0; i--) {
        out[i] = (uint8_t)(n & 0xFF);
        n >>= 8;
    }
}

/*
 * The remaining exports are "internal, exposed for testing" in the reference (bytes.h:66-74,
 * eip4844.h:84, eip7594.h:58-68); its Go and Rust test-suites call them.  fr_t / g1_t are private to the
 * library: here an fr_t carries the canonical big-endian scalar and a g1_t the
 * validated 48-byte compression, so the conversions are copies; validation and the mod-r reduction
 * still run on the GPU.
 */
static const uint8_t BLS_MODULUS_BE[32] = {0x73, 0xed, 0xa7, 0x53, 0x29, 0x9d, 0x7d, 0x48, 0x33, 0x39, 0xd8, 0x08, 0x09, 0xa1, 0xd8, 0x05,
                                           0x53, 0xbd, 0xa4, 0x02, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0xff, 0xff, 0xff, 0x00, 0x00, 0x00, 0x01};

C_KZG_RET bytes_to_bls_field(fr_t *out, const Bytes32 *b) { /* bytes.c:64: canonical (< r) or BADARGS */
    if (memcmp(b->bytes, BLS_MODULUS_BE, 32) >= 0) return C_KZG_BADARGS;
    memcpy(out, b->bytes, 32);
    return C_KZG_OK;
}

void bytes_from_bls_field(Bytes32 *out, const fr_t *in) { memcpy(out->bytes, in, 32); } /* bytes.c:52 */

static C_KZG_RET bytes_to_g1(g1_t *out, const Bytes48 *b) { /* validate_kzg_g1, bytes.c:81 */
    int ok = 0;
    int rc = ckzg_b200_validate_g1(&ok, b->bytes);
    if (rc != 0) return (C_KZG_RET)rc;
    if (!ok) return C_KZG_BADARGS;
    memset(out, 0, sizeof(*out));
    memcpy(out, b->bytes, 48);
    return C_KZG_OK;
}
C_KZG_RET bytes_to_kzg_commitment(g1_t *out, const Bytes48 *b) { return bytes_to_g1(out, b); } /* bytes.c:101 */
C_KZG_RET bytes_to_kzg_proof(g1_t *out, const Bytes48 *b) { return bytes_to_g1(out, b); }      /* bytes.c:112 */
void bytes_from_g1(Bytes48 *out, const g1_t *in) { memcpy(out->bytes, in, 48); }               /* bytes.c:42 */

void compute_challenge(fr_t *eval_challenge_out, const Blob *blob, const g1_t *commitment) { /* eip4844.c:147 */
    uint8_t z[32];
    /* the reference's signature has no error channel: a failed device call must not come back as a zero
     * challenge that looks like a result */
    int rc = ckzg_b200_compute_challenge(NULL, z, blob->bytes, (const uint8_t *)commitment);
    if (rc != 0) {
        fprintf(stderr, "ckzg_b200: compute_challenge failed on the device (code %d); no CPU fallback exists\n", rc);
        abort();
    }
    memset(eval_challenge_out, 0, sizeof(*eval_challenge_out));
    memcpy(eval_challenge_out, z, 32);
}

C_KZG_RET compute_verify_cell_kzg_proof_batch_challenge( /* eip7594.c:390-482 */
    fr_t *challenge_out,
    const Bytes48 *commitments_bytes,
    uint64_t num_commitments,
    const uint64_t *commitment_indices,
    const uint64_t *cell_indices,
    const Cell *cells,
    const Bytes48 *proofs_bytes,
    uint64_t num_cells
) {
    ckzg_host_sha256 h;
    uint8_t u64[8], digest[32], out[32];
    ckzg_host_sha256_init(&h);
    ckzg_host_sha256_update(&h, "RCKZGCBATCH__V1_", 16);
    bytes_from_uint64(u64, FIELD_ELEMENTS_PER_BLOB);
    ckzg_host_sha256_update(&h, u64, 8);
    bytes_from_uint64(u64, FIELD_ELEMENTS_PER_CELL);
    ckzg_host_sha256_update(&h, u64, 8);
    bytes_from_uint64(u64, num_commitments);
    ckzg_host_sha256_update(&h, u64, 8);
    bytes_from_uint64(u64, num_cells);
    ckzg_host_sha256_update(&h, u64, 8);
    ckzg_host_sha256_update(&h, commitments_bytes, (size_t)num_commitments * BYTES_PER_COMMITMENT);
    for (uint64_t i = 0; i < num_cells; i++) {
        bytes_from_uint64(u64, commitment_indices[i]);
        ckzg_host_sha256_update(&h, u64, 8);
        bytes_from_uint64(u64, cell_indices[i]);
        ckzg_host_sha256_update(&h, u64, 8);
        ckzg_host_sha256_update(&h, cells[i].bytes, BYTES_PER_CELL);
        ckzg_host_sha256_update(&h, proofs_bytes[i].bytes, BYTES_PER_PROOF);
    }
    ckzg_host_sha256_final(&h, digest);
    int rc = ckzg_b200_hash_to_bls_field(out, digest);
    if (rc != 0) return (C_KZG_RET)rc;
    memcpy(challenge_out, out, 32);
    return C_KZG_OK;
}
