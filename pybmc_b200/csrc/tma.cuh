// mbarrier + 1-D TMA bulk-copy helpers (cp.async.bulk; UBLKCP in SASS) used to stage sample and
// design-matrix tiles in shared memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bmc {

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
                 "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(addr),
        "r"(phase)
        : "memory");
}
// The same wait for a thread that has nothing else to do (a producer lane): back off between polls so the
// spin does not take issue slots from the warps doing the arithmetic on the same scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t phase, uint32_t ns = 64) {
    const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(phase)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
// 1-D bulk copy global -> shared through the TMA unit (UBLKCP in SASS); bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
        "l"(gmem_src), "r"(bytes), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
        : "memory");
}

}  // namespace bmc
