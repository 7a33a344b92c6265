/*
 * host_sha256.c -- SHA-256 on the host for the BATCH transcript challenges only.
 *
 * The batch verifiers hash one serial transcript per call: 32 + 160 n bytes for
 * verify_blob_kzg_proof_batch (src/eip4844/eip4844.c:597-680).  A hash chain cannot be parallelised,
 * and one GPU thread needs ~2.8 us per 64-byte block (measured: 28.7 ms for n = 4096, 41 % of the whole
 * verification) where a host core with SHA-NI needs ~0.04 us.  SURVEY.md section 7.3-5 / 8(e) names exactly
 * this split: the per-blob 131 KB hashes (n independent chains) stay on the GPU, the tiny serial
 * transcript hash is host C, between two device stages.  It is not a fallback: there is no device
 * implementation of this step to fall back from, and no field or curve arithmetic happens here (the
 * digest is reduced mod r on the device).
 *
 * FIPS 180-4.  Replaces blst_sha256 (blst/src/exports.c:558) for that one call site.
 */
#include "host_sha256.h"

#include <string.h>

/* CKZG_HOST_SHA_PORTABLE forces the portable rounds (tests/test_host_sha256.py checks both forms) */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(CKZG_HOST_SHA_PORTABLE)
#include <immintrin.h>
#define HAVE_X86_SHA 1
#else
#define HAVE_X86_SHA 0
#endif

static const uint32_t K[64] __attribute__((aligned(16))) = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void blocks_portable(uint32_t h[8], const uint8_t *p, size_t nblocks) {
    uint32_t w[64];
    while (nblocks--) {
        for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        p += 64;
    }
}

#if HAVE_X86_SHA
/* Four rounds: the two sha256rnds2 of one 128-bit group of message words + constants. */
#define SHA_RNDS4(MSG, KIDX)                                                            \
    do {                                                                                \
        __m128i wk_ = _mm_add_epi32((MSG), _mm_load_si128((const __m128i *)&K[KIDX])); \
        s1 = _mm_sha256rnds2_epu32(s1, s0, wk_);                                        \
        wk_ = _mm_shuffle_epi32(wk_, 0x0E);                                             \
        s0 = _mm_sha256rnds2_epu32(s0, s1, wk_);                                        \
    } while (0)
/* Next four schedule words from the previous sixteen: M0 <- f(M0, M1, M2, M3). */
#define SHA_SCHED(M0, M1, M2, M3) M0 = _mm_sha256msg2_epu32(_mm_add_epi32(_mm_sha256msg1_epu32(M0, M1), _mm_alignr_epi8(M3, M2, 4)), M3)

/* Fully unrolled, message words in four named registers (the rolled form kept them in an indexed array on the
 * stack: 15-20 % slower).  The serial transcript hash of verify_cell_kzg_proof_batch is 2112 bytes per cell --
 * 69 MB for 256 blobs x 128 cells -- and is the largest single item of that call. */
__attribute__((target("sha,sse4.1,ssse3"))) static void blocks_shani(uint32_t h[8], const uint8_t *p, size_t nblocks) {
    const __m128i shuf = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i t = _mm_loadu_si128((const __m128i *)&h[0]);   /* a b c d */
    __m128i s1 = _mm_loadu_si128((const __m128i *)&h[4]);  /* e f g h */
    t = _mm_shuffle_epi32(t, 0xB1);                         /* c d a b */
    s1 = _mm_shuffle_epi32(s1, 0x1B);                       /* h g f e */
    __m128i s0 = _mm_alignr_epi8(t, s1, 8);                 /* a b e f */
    s1 = _mm_blend_epi16(s1, t, 0xF0);                      /* c d g h */
    while (nblocks--) {
        const __m128i save0 = s0, save1 = s1;
        __m128i m0 = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 0)), shuf);
        __m128i m1 = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 16)), shuf);
        __m128i m2 = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 32)), shuf);
        __m128i m3 = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 48)), shuf);
        SHA_RNDS4(m0, 0);  SHA_SCHED(m0, m1, m2, m3);
        SHA_RNDS4(m1, 4);  SHA_SCHED(m1, m2, m3, m0);
        SHA_RNDS4(m2, 8);  SHA_SCHED(m2, m3, m0, m1);
        SHA_RNDS4(m3, 12); SHA_SCHED(m3, m0, m1, m2);
        SHA_RNDS4(m0, 16); SHA_SCHED(m0, m1, m2, m3);
        SHA_RNDS4(m1, 20); SHA_SCHED(m1, m2, m3, m0);
        SHA_RNDS4(m2, 24); SHA_SCHED(m2, m3, m0, m1);
        SHA_RNDS4(m3, 28); SHA_SCHED(m3, m0, m1, m2);
        SHA_RNDS4(m0, 32); SHA_SCHED(m0, m1, m2, m3);
        SHA_RNDS4(m1, 36); SHA_SCHED(m1, m2, m3, m0);
        SHA_RNDS4(m2, 40); SHA_SCHED(m2, m3, m0, m1);
        SHA_RNDS4(m3, 44); SHA_SCHED(m3, m0, m1, m2);
        SHA_RNDS4(m0, 48);
        SHA_RNDS4(m1, 52);
        SHA_RNDS4(m2, 56);
        SHA_RNDS4(m3, 60);
        s0 = _mm_add_epi32(s0, save0);
        s1 = _mm_add_epi32(s1, save1);
        p += 64;
    }
    t = _mm_shuffle_epi32(s0, 0x1B);   /* f e b a */
    s1 = _mm_shuffle_epi32(s1, 0xB1);  /* d c h g */
    s0 = _mm_blend_epi16(t, s1, 0xF0); /* d c b a */
    s1 = _mm_alignr_epi8(s1, t, 8);    /* h g f e */
    _mm_storeu_si128((__m128i *)&h[0], s0);
    _mm_storeu_si128((__m128i *)&h[4], s1);
}
#undef SHA_RNDS4
#undef SHA_SCHED
#endif

static void blocks(uint32_t h[8], const uint8_t *p, size_t n) {
#if HAVE_X86_SHA
    static int have = -1;
    if (have < 0) have = __builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3");
    if (have) {
        blocks_shani(h, p, n);
        return;
    }
#endif
    blocks_portable(h, p, n);
}

void ckzg_host_sha256_init(ckzg_host_sha256 *s) {
    static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    memcpy(s->h, iv, sizeof(iv));
    s->fill = 0;
    s->total = 0;
}

void ckzg_host_sha256_update(ckzg_host_sha256 *s, const void *data, size_t n) {
    const uint8_t *p = (const uint8_t *)data;
    s->total += n;
    if (s->fill) {
        size_t take = 64 - s->fill;
        if (take > n) take = n;
        memcpy(s->buf + s->fill, p, take);
        s->fill += take;
        p += take;
        n -= take;
        if (s->fill == 64) {
            blocks(s->h, s->buf, 1);
            s->fill = 0;
        }
    }
    if (n >= 64) {
        size_t nb = n / 64;
        blocks(s->h, p, nb);
        p += 64 * nb;
        n -= 64 * nb;
    }
    if (n) {
        memcpy(s->buf, p, n);
        s->fill = n;
    }
}

void ckzg_host_sha256_final(ckzg_host_sha256 *s, uint8_t out[32]) {
    uint64_t bits = s->total * 8;
    uint8_t pad[72] = {0x80};
    size_t padlen = (s->fill < 56) ? (56 - s->fill) : (120 - s->fill);
    uint8_t len[8];
    for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i));
    ckzg_host_sha256_update(s, pad, padlen);
    ckzg_host_sha256_update(s, len, 8);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(s->h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(s->h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(s->h[i] >> 8);
        out[4 * i + 3] = (uint8_t)s->h[i];
    }
}
