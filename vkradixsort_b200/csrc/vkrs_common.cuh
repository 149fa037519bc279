// vkrs_common.cuh -- small device helpers shared by every kernel of the sort.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vkrs {

constexpr int RADIX_BITS = 8;        // 8 bits per pass (multi_radixsort.comp:12, RADIX_SORT_BINS 256)
constexpr int RADIX = 1 << RADIX_BITS;

// Chained-scan tile status word: [31:30] flag, [29:0] count.  30 bits of count is the
// reference's own limit (its byte sizes are uint32, MultiRadixSort.h:29-31 => N < 2^30).
constexpr uint32_t STATUS_FLAG_AGGREGATE = 1u << 30; // this tile's own digit count
constexpr uint32_t STATUS_FLAG_INCLUSIVE = 2u << 30; // count of this tile and all tiles before it
constexpr uint32_t STATUS_FLAG_MASK = 3u << 30;
constexpr uint32_t STATUS_VALUE_MASK = ~STATUS_FLAG_MASK;
constexpr uint32_t LOOKBACK_SPIN_LIMIT = 1u << 24; // polls before a tile gives up and raises the error flag

enum DeviceError : uint32_t { DEVERR_NONE = 0, DEVERR_LOOKBACK_TIMEOUT = 1 };

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// L2-coherent (never L1-cached) accesses for the tile status words that other CTAs poll.
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Streaming loads/stores for key data that is touched once per pass: keep it out of L1.
template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
    return __ldcs(p);
}

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT key, uint32_t shift) {
    return static_cast<uint32_t>(key >> shift) & (RADIX - 1);
}

// Peer mask of the lanes holding the same 8-bit digit.
//   MATCH_BALLOT: 8 ballots + LOP3s (no shared memory, no special unit)
//   MATCH_HW    : the match.any instruction
enum MatchMode { MATCH_BALLOT = 0, MATCH_HW = 1 };

template <int MODE>
__device__ __forceinline__ uint32_t match_digit(uint32_t digit) {
    if (MODE == MATCH_HW) {
        return __match_any_sync(0xffffffffu, digit);
    } else {
        uint32_t mask = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < RADIX_BITS; ++b) {
            const bool bit = (digit >> b) & 1u;
            const uint32_t vote = __ballot_sync(0xffffffffu, bit);
            mask &= bit ? vote : ~vote;
        }
        return mask;
    }
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread over the first 256 threads of the block (one per
// digit).  Every thread of the block must call it (it contains __syncthreads); threads
// >= 256 pass 0 and ignore the result.  `scratch` = 8 uint32 of shared memory.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *scratch, uint32_t *total_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t incl = warp_inclusive_scan(v, lane);
    if (warp < 8 && lane == 31) scratch[warp] = incl;
    __syncthreads();
    uint32_t warp_prefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s = scratch[w];
        if (w < warp) warp_prefix += s;
        total += s;
    }
    if (total_out) *total_out = total;
    return warp_prefix + incl - v;
}

} // namespace vkrs
