// Bulk asynchronous copies global -> shared memory (the TMA unit's 1-D form: cp.async.bulk, SASS UBLKCP) with
// mbarrier completion -- sm_90+ PTX, written for sm_100a.  One thread arms the barrier with the byte count and
// issues the copy; the copy engine moves the bytes without occupying registers or issue slots, and every thread
// that needs the data waits on the barrier's phase.
#pragma once
#include <stdint.h>

namespace kzg {

__device__ __forceinline__ uint32_t tma_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_addr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy before any copy names it
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_addr(bar)), "r"(bytes) : "memory");
}
// bytes: multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tma_smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes),
                 "r"(tma_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KZG_TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KZG_TMA_DONE;\n"
        "bra KZG_TMA_WAIT;\n"
        "KZG_TMA_DONE:\n"
        "}\n" ::"r"(tma_smem_addr(bar)),
        "r"(phase)
        : "memory");
}

}  // namespace kzg
