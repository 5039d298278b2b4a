// tr_tma.cuh -- bulk asynchronous copies global -> shared memory through the TMA unit (cp.async.bulk, PTX ISA 8.0+;
// SASS: UBLKCP) with mbarrier completion. One thread issues the copy; every thread that reads the data waits on the
// barrier's phase. Source, destination and size must be multiples of 16 bytes.
#pragma once
#include <cstdint>

namespace trd {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Call from ONE thread, then __syncthreads() before anyone arrives or waits.
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(mbar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Call from ONE thread: arms the barrier with the byte count and starts the copy.
__device__ __forceinline__ void tma_load_bulk(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(mbar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TR_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra TR_MBAR_WAIT;\n\t}"
        :: "r"(smem_addr(mbar)), "r"(phase) : "memory");
}

} // namespace trd
