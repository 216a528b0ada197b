// snp_tma.cuh -- 1-D TMA bulk copies + mbarrier primitives (sm_100a PTX; SASS: UBLKCP, SYNCS.*).
//
// cp.async.bulk needs no tensor map for 1-D copies: 16-byte-aligned source, destination and size.
// Under SNP_EMU (tests/cpp/simt_emu.h) the copies are synchronous memcpy and the barriers no-ops, so
// the host emulator checks the ring bookkeeping around them, not the asynchrony.
#pragma once
#include "snp_common.cuh"

namespace snp {

#ifdef SNP_EMU
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *, uint32_t) {}
__device__ __forceinline__ void mbar_arrive(uint64_t *) {}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *, uint32_t) { return true; }
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *) {
    memcpy(smem_dst, gmem_src, bytes);
}
__device__ __forceinline__ void mbar_fence_init() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

}  // namespace snp
