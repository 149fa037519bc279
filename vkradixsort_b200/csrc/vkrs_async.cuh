// vkrs_async.cuh -- sm_100a asynchronous-copy and barrier primitives used by the pipelined
// digit pass: mbarrier (phase-tracked producer/consumer signals in shared memory), the 1-D TMA
// bulk copy global -> shared (cp.async.bulk, completes on an mbarrier), named barriers for
// sub-CTA groups.  Thin wrappers over the PTX, nothing else.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vkrs {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// Makes freshly initialised mbarriers visible to the async (TMA) proxy.
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Blocks until the phase with the given parity has completed.  try_wait suspends the thread in
// hardware for a bounded time per attempt.  A wait still unsatisfied after 10 s of wall time can
// only be a protocol bug or a lost peer, so it traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t_start = 0;
    for (uint32_t attempt = 0;; ++attempt) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if ((attempt & 255u) == 255u) {
            const unsigned long long now = global_timer_ns();
            if (t_start == 0) t_start = now;
            else if (now - t_start > 10000000000ull) __trap();
        }
    }
}

// Same, for threads that have nothing else to do and share their scheduler with busy warps: one
// probe, then sleep `ns` before the next, so the waiting warp stays out of the issue slots.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity, uint32_t ns) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t_start = 0;
    for (uint32_t attempt = 0;; ++attempt) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        __nanosleep(ns);
        if ((attempt & 1023u) == 1023u) {
            const unsigned long long now = global_timer_ns();
            if (t_start == 0) t_start = now;
            else if (now - t_start > 10000000000ull) __trap();
        }
    }
}

// 1-D TMA bulk copy: `bytes` (multiple of 16) from 16-byte aligned global memory to 16-byte
// aligned shared memory; the mbarrier receives complete_tx(bytes) when the data has landed.
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 1-D TMA bulk copy shared -> global (bulk async-group completion).  The shared-memory source must have been written
// before a fence_proxy_async() of the writing threads and a barrier; the issuing thread commits the group and, before
// the source is overwritten, waits for the group's reads (bulk_store_wait_read) -- before the kernel ends, for the
// group itself (bulk_store_wait_all).
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Named barrier over a subset of the CTA's warps (id 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

} // namespace vkrs
