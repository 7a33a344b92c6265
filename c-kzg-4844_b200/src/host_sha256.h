/* Host SHA-256 for the serial batch-transcript challenges (see host_sha256.c for the rationale). */
#ifndef CKZG_HOST_SHA256_H
#define CKZG_HOST_SHA256_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    uint32_t h[8];
    uint8_t buf[64];
    size_t fill;
    uint64_t total;
} ckzg_host_sha256;
void ckzg_host_sha256_init(ckzg_host_sha256 *s);
void ckzg_host_sha256_update(ckzg_host_sha256 *s, const void *data, size_t n);
void ckzg_host_sha256_final(ckzg_host_sha256 *s, uint8_t out[32]);
#ifdef __cplusplus
}
#endif
#endif
