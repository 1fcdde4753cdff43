// mbarrier / bulk-copy / descriptor helpers shared by the tcgen05 contraction
// (gt_contract_tc2.cuh) and the streaming correction kernel (gt_corr_stream.cuh).
#pragma once

#include "gt_common.cuh"

namespace gtk {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: never hangs the GPU; gives up after ~0.2 s or as soon as
// another thread of the CTA has flagged a failure
// BACKOFF: the single-lane producer / MMA warps sleep between polls so that
// their spinning does not take issue slots from the expansion warps.
template <bool BACKOFF = false>
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, volatile int *fail) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 23); ++spin) {
        if ((spin & 1023) == 1023 && *fail) return false;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
        if (BACKOFF) __nanosleep(64);
    }
    return false;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 ldg_stream16(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// shared-memory matrix descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address  [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused, 1)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}

}  // namespace gtk
