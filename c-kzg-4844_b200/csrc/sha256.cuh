// SHA-256 for the Fiat-Shamir challenges (replaces blst_sha256, blst/src/exports.c:558 /
// blst/src/sha256.h, as used by src/eip4844/eip4844.c:176,667 and src/eip7594/eip7594.c:474).
// FIPS 180-4, streaming over 64-byte blocks supplied as 16 big-endian words.
#pragma once
#include "field.cuh"

namespace kzg {

KZG_CONST uint32_t SHA256_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

struct Sha256 {
    uint32_t h[8];
};

KZG_HD uint32_t sha_rotr(uint32_t x, int n) {
#if KZG_DEVICE_PATH
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

KZG_HD void sha256_init(Sha256& s) {
    s.h[0] = 0x6a09e667u; s.h[1] = 0xbb67ae85u; s.h[2] = 0x3c6ef372u; s.h[3] = 0xa54ff53au;
    s.h[4] = 0x510e527fu; s.h[5] = 0x9b05688cu; s.h[6] = 0x1f83d9abu; s.h[7] = 0x5be0cd19u;
}

// one compression; w[16] = the block as big-endian words (clobbered)
KZG_HD void sha256_block(Sha256& s, uint32_t* w) {
    uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
    for (int t = 0; t < 64; t++) {
        if (t >= 16) {
            uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
        }
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + SHA256_K[t] + w[t & 15];
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1;
        d = c; c = b; b = a; a = t1 + t2;
    }
    s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d;
    s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

KZG_HD uint32_t be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

// Byte-stream front end (small inputs / tails).  buf holds a partial block.
struct Sha256Stream {
    Sha256 st;
    uint8_t buf[64];
    uint32_t fill;
    uint64_t total;
};
KZG_HD void sha256_stream_init(Sha256Stream& s) {
    sha256_init(s.st);
    s.fill = 0;
    s.total = 0;
}
KZG_HD void sha256_stream_flush(Sha256Stream& s) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = be32(s.buf + 4 * i);
    sha256_block(s.st, w);
    s.fill = 0;
}
KZG_HD void sha256_stream_update(Sha256Stream& s, const uint8_t* p, uint64_t n) {
    s.total += n;
    for (uint64_t i = 0; i < n; i++) {
        s.buf[s.fill++] = p[i];
        if (s.fill == 64) sha256_stream_flush(s);
    }
}
KZG_HD void sha256_stream_final(Sha256Stream& s, uint8_t out[32]) {
    uint64_t bits = s.total * 8;
    s.buf[s.fill++] = 0x80;
    if (s.fill > 56) {
        while (s.fill < 64) s.buf[s.fill++] = 0;
        sha256_stream_flush(s);
    }
    while (s.fill < 56) s.buf[s.fill++] = 0;
    for (int i = 0; i < 8; i++) s.buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
    s.fill = 64;
    sha256_stream_flush(s);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(s.st.h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(s.st.h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(s.st.h[i] >> 8);
        out[4 * i + 3] = (uint8_t)s.st.h[i];
    }
}

}  // namespace kzg
