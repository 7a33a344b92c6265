/*
 * A consumer of the frozen c-kzg-4844 C API, written the way the reference's own C users are
 * (src/test/tests.c:2268-2290 loads the setup the same way): nothing here knows about CUDA.
 *
 *   gcc -Iinclude examples/c_consumer.c -Lc-kzg-4844_b200 -lckzg_b200 -Wl,-rpath,$PWD/c-kzg-4844_b200 -o c_consumer
 *   (or: gcc examples/c_consumer.c $(pkg-config --cflags --libs c-kzg-4844_b200/ckzg_b200.pc) -o c_consumer)
 *   ./c_consumer c-kzg-4844_b200/data/trusted_setup.txt [batch]
 *
 * With a batch size (e.g. 640) it also verifies ONE verify_blob_kzg_proof_batch / verify_cell_kzg_proof_batch over
 * that many distinct blobs and a negative control for each.  Nothing changes here when the library drives several
 * GPUs: CKZG_B200_DEVICES=0,1 ./c_consumer ... 640 shards those calls inside the library (include/ckzg_b200.h).
 *
 * Exit code 0 = commitment, proof, blob verification, cells + cell proofs, recovery and cell verification
 * all agree with each other; 2 = no usable CUDA device (load_trusted_setup returned C_KZG_ERROR: there is no
 * CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ckzg.h"

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "c-kzg-4844_b200/data/trusted_setup.txt";
    FILE *f = fopen(path, "r");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", path);
        return 1;
    }
    KZGSettings s;
    C_KZG_RET rc = load_trusted_setup_file(&s, f, 0);
    fclose(f);
    if (rc == C_KZG_ERROR) {
        fprintf(stderr, "load_trusted_setup_file: C_KZG_ERROR (no CUDA device?)\n");
        return 2;
    }
    if (rc != C_KZG_OK) return 1;

    Blob *blob = calloc(1, sizeof(Blob));
    Cell *cells = calloc(CELLS_PER_EXT_BLOB, sizeof(Cell)), *rec = calloc(CELLS_PER_EXT_BLOB, sizeof(Cell));
    KZGProof *cproofs = calloc(CELLS_PER_EXT_BLOB, sizeof(KZGProof)), *rproofs = calloc(CELLS_PER_EXT_BLOB, sizeof(KZGProof));
    if (!blob || !cells || !rec || !cproofs || !rproofs) return 1;
    for (size_t i = 0; i < FIELD_ELEMENTS_PER_BLOB; i++) { /* small canonical field elements */
        blob->bytes[32 * i + 30] = (uint8_t)(i >> 8);
        blob->bytes[32 * i + 31] = (uint8_t)i;
        blob->bytes[32 * i + 17] = (uint8_t)(i * 7 + 3);
    }
    KZGCommitment c;
    KZGProof p;
    bool ok = false;
    int fail = 0;
    fail |= blob_to_kzg_commitment(&c, blob, &s) != C_KZG_OK;
    fail |= compute_blob_kzg_proof(&p, blob, &c, &s) != C_KZG_OK;
    fail |= verify_blob_kzg_proof(&ok, blob, &c, &p, &s) != C_KZG_OK || !ok;
    fail |= verify_blob_kzg_proof_batch(&ok, blob, &c, &p, 1, &s) != C_KZG_OK || !ok;
    p.bytes[47] ^= 1; /* a damaged proof is either an invalid point (BADARGS) or fails the check */
    rc = verify_blob_kzg_proof(&ok, blob, &c, &p, &s);
    fail |= !(rc == C_KZG_BADARGS || (rc == C_KZG_OK && !ok));

    fail |= compute_cells_and_kzg_proofs(cells, cproofs, blob, &s) != C_KZG_OK;
    fail |= memcmp(cells, blob->bytes, BYTES_PER_BLOB) != 0; /* the first half of the extension is the blob */
    uint64_t idx[CELLS_PER_EXT_BLOB];
    Cell *half = calloc(CELLS_PER_BLOB, sizeof(Cell));
    Bytes48 *cms = calloc(CELLS_PER_EXT_BLOB, sizeof(Bytes48));
    if (!half || !cms) return 1;
    for (size_t k = 0; k < CELLS_PER_BLOB; k++) {
        idx[k] = 2 * k + 1;
        memcpy(&half[k], &cells[2 * k + 1], sizeof(Cell));
    }
    fail |= recover_cells_and_kzg_proofs(rec, rproofs, idx, half, CELLS_PER_BLOB, &s) != C_KZG_OK;
    fail |= memcmp(rec, cells, CELLS_PER_EXT_BLOB * sizeof(Cell)) != 0;
    fail |= memcmp(rproofs, cproofs, CELLS_PER_EXT_BLOB * sizeof(KZGProof)) != 0;
    for (size_t k = 0; k < CELLS_PER_EXT_BLOB; k++) {
        idx[k] = k;
        memcpy(&cms[k], &c, sizeof(Bytes48));
    }
    fail |= verify_cell_kzg_proof_batch(&ok, cms, idx, cells, cproofs, CELLS_PER_EXT_BLOB, &s) != C_KZG_OK || !ok;
    idx[5] = 6; /* wrong index for cell 5 */
    fail |= verify_cell_kzg_proof_batch(&ok, cms, idx, cells, cproofs, CELLS_PER_EXT_BLOB, &s) != C_KZG_OK || ok;

    /* ---- optional batch phase: n distinct blobs through the batch verifiers ---- */
    const size_t n = argc > 2 ? (size_t)atol(argv[2]) : 0;
    if (n > 1) {
        Blob *blobs = calloc(n, sizeof(Blob));
        Bytes48 *bc = calloc(n, sizeof(Bytes48)), *bp = calloc(n, sizeof(Bytes48));
        if (!blobs || !bc || !bp) return 1;
        for (size_t b = 0; b < n; b++) {
            for (size_t i = 0; i < FIELD_ELEMENTS_PER_BLOB; i++) {
                uint8_t *e = blobs[b].bytes + 32 * i;
                uint32_t v = (uint32_t)(b * 2654435761u + i * 40503u + 12345u);
                e[1] = (uint8_t)(v >> 24); /* top byte stays 0: canonical */
                e[9] = (uint8_t)(v >> 16);
                e[20] = (uint8_t)(v >> 8);
                e[31] = (uint8_t)v;
                e[30] = (uint8_t)b;
                e[29] = (uint8_t)(b >> 8);
            }
            fail |= blob_to_kzg_commitment((KZGCommitment *)&bc[b], &blobs[b], &s) != C_KZG_OK;
            fail |= compute_blob_kzg_proof((KZGProof *)&bp[b], &blobs[b], &bc[b], &s) != C_KZG_OK;
        }
        fail |= verify_blob_kzg_proof_batch(&ok, blobs, bc, bp, n, &s) != C_KZG_OK || !ok;
        Bytes48 t = bp[n - 2]; /* proofs of the last two blobs swapped: valid points, wrong proofs */
        bp[n - 2] = bp[n - 1];
        bp[n - 1] = t;
        fail |= verify_blob_kzg_proof_batch(&ok, blobs, bc, bp, n, &s) != C_KZG_OK || ok;
        bp[n - 1] = bp[n - 2];
        bp[n - 2] = t;
        blobs[n / 2].bytes[32 * 77] = 0xff; /* one non-canonical field element */
        fail |= verify_blob_kzg_proof_batch(&ok, blobs, bc, bp, n, &s) != C_KZG_BADARGS;
        blobs[n / 2].bytes[32 * 77] = 0;
        /* cells of the first min(n, 64) blobs, all 128 cells each, in one verify_cell_kzg_proof_batch */
        const size_t nb = n < 64 ? n : 64, nc = nb * CELLS_PER_EXT_BLOB;
        Cell *ac = calloc(nc, sizeof(Cell));
        KZGProof *ap = calloc(nc, sizeof(KZGProof));
        Bytes48 *acm = calloc(nc, sizeof(Bytes48));
        uint64_t *aidx = calloc(nc, sizeof(uint64_t));
        if (!ac || !ap || !acm || !aidx) return 1;
        for (size_t b = 0; b < nb; b++) {
            fail |= compute_cells_and_kzg_proofs(ac + b * CELLS_PER_EXT_BLOB, ap + b * CELLS_PER_EXT_BLOB, &blobs[b], &s) != C_KZG_OK;
            for (size_t k = 0; k < CELLS_PER_EXT_BLOB; k++) {
                aidx[b * CELLS_PER_EXT_BLOB + k] = k;
                acm[b * CELLS_PER_EXT_BLOB + k] = bc[b];
            }
        }
        fail |= verify_cell_kzg_proof_batch(&ok, acm, aidx, ac, ap, nc, &s) != C_KZG_OK || !ok;
        ac[nc - 1].bytes[2047] ^= 1; /* last cell of the last blob damaged */
        fail |= verify_cell_kzg_proof_batch(&ok, acm, aidx, ac, ap, nc, &s) != C_KZG_OK || ok;
        printf("c_consumer: batch phase n=%zu (%zu cells) %s\n", n, nc, fail ? "MISMATCH" : "ok");
    }

    free_trusted_setup(&s);
    free_trusted_setup(&s); /* safe to call twice (setup.c:162-190) */
    printf(fail ? "c_consumer: MISMATCH\n" : "c_consumer: ok\n");
    return fail ? 1 : 0;
}
