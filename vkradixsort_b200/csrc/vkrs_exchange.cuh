// vkrs_exchange.cuh -- the multi-GPU bucket exchange's control work, on the device: the plan (which rank owns which
// contiguous range of the 256 buckets, where this rank's part of every bucket lands in the owner's receive buffer)
// from the all-gathered bucket counts, and a barrier between the ranks of one node through flags in peer memory.
// With these the exchange step never returns to the host: all-gather of the counts -> exchange_plan_kernel ->
// segmented_scatter_kernel<P2P> (stores cross NVLink as tiles leave the SM) -> peer_signal_wait_kernel -> local sort.
// The reference is single-device; BASELINE.json configs[4] defines this extension.
#pragma once
#include "vkrs_common.cuh"

namespace vkrs {

constexpr int EXCHANGE_MAX_RANKS = 64;

// boundaries[r], r = 0..world: rank r owns buckets [boundaries[r], boundaries[r+1]).  Rank r's range ends at the
// bucket where the running total is closest to r/world of all keys (integer arithmetic, identical to
// vkradixsort_b200/dist.py:plan_exchange, which the host evaluates on the same counts for its own bookkeeping).
//   counts      [world][256] uint32, all-gathered (row s = bucket counts of rank s)
//   peer_keys / peer_vals  [world] device addresses of the ranks' receive buffers (peer_vals may be NULL)
//   dst_tables  out, 4 x 256 uint64: address of this rank's part of bucket b in the owner's key buffer | same for
//               payloads | first bucket of the owner's range | end bucket of the owner's range
//   summary     out, 4 + world uint32: keys this rank receives | largest receive count of any rank | gate (1 = the largest
//               range fits `capacity` keys and, if max_imbalance_permille != 0, largest / mean <= permille / 1000) | 0 |
//               boundaries[1..world-1]
__global__ void __launch_bounds__(RADIX)
exchange_plan_kernel(const uint32_t *__restrict__ counts, uint32_t world, uint32_t rank, const unsigned long long *__restrict__ peer_keys,
                     const unsigned long long *__restrict__ peer_vals, unsigned long long *__restrict__ dst_tables,
                     uint32_t *__restrict__ summary, uint32_t capacity, uint32_t max_imbalance_permille) {
    __shared__ unsigned long long cum[RADIX];       // inclusive running total over the buckets, all ranks
    __shared__ uint32_t bounds[EXCHANGE_MAX_RANKS + 1];
    __shared__ unsigned long long part[EXCHANGE_MAX_RANKS + 1];
    __shared__ uint32_t scratch[8];
    const uint32_t b = threadIdx.x;
    uint32_t total_b = 0, low = 0;
    for (uint32_t s = 0; s < world; ++s) {
        const uint32_t c = counts[s * RADIX + b];
        if (s < rank) low += c;
        total_b += c;
    }

    // inclusive scan of the bucket totals (64-bit: up to 64 x 2^30 keys)
    {
        const int lane = b & 31, warp = b >> 5;
        unsigned long long v = total_b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        __shared__ unsigned long long wsum[8];
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        unsigned long long pre = 0;
        for (int w = 0; w < warp; ++w) pre += wsum[w];
        cum[b] = v + pre;
    }
    __syncthreads();
    const unsigned long long total = cum[RADIX - 1];
    // boundary of rank r (1 <= r < world): idx = first bucket whose running total reaches total * r / world; cut before
    // or after it, whichever lands closer; monotone, empty ranges allowed.  One thread: world <= 64 short searches.
    if (b == 0) {
        bounds[0] = 0;
        for (uint32_t r = 1; r < world; ++r) {
            uint32_t cut = 0;
            if (total > 0) {
                const unsigned long long tr = total * r; // target * world
                uint32_t lo = 0, hi = RADIX;             // first idx with cum[idx] * world >= tr
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (cum[mid] * world >= tr) hi = mid;
                    else lo = mid + 1;
                }
                uint32_t idx = lo < (uint32_t) RADIX - 1 ? lo : (uint32_t) RADIX - 1;
                const unsigned long long before = idx > 0 ? cum[idx - 1] : 0ull;
                // (target - before) <= (cum[idx] - target), times world
                const long long d0 = (long long) tr - (long long) (before * world), d1 = (long long) (cum[idx] * world) - (long long) tr;
                cut = d0 <= d1 ? idx : idx + 1;
            }
            const uint32_t prev = bounds[r - 1];
            cut = cut < prev ? prev : cut;
            bounds[r] = cut > (uint32_t) RADIX ? (uint32_t) RADIX : cut;
        }
        bounds[world] = RADIX;
    }
    __syncthreads();
    // owner of bucket b, and where this rank's keys of the owner's whole range start in the owner's receive buffer:
    // the buffer is laid out source rank by source rank, so the offset is the sum over lower ranks of their keys in
    // the owner's range -- the same for every bucket of the range (the scatter kernel writes one run per owner and tile)
    uint32_t owner = 0;
    while (owner + 1 < world && b >= bounds[owner + 1]) ++owner;
    if (b <= world) part[b] = 0;
    __syncthreads();
    atomicAdd(&part[owner], (unsigned long long) low);
    __syncthreads();
    const unsigned long long off = part[owner];
    dst_tables[b] = peer_keys[owner] + 4ull * off;
    dst_tables[RADIX + b] = peer_vals ? peer_vals[owner] + 4ull * off : 0ull;
    dst_tables[2 * RADIX + b] = bounds[owner];
    dst_tables[3 * RADIX + b] = bounds[owner + 1];
    // summary: what this rank receives, the largest receive count of any rank (the buffers must hold it)
    if (b == 0) scratch[0] = 0;
    __syncthreads();
    if (b < world) {
        const uint32_t r = b;
        const unsigned long long hi_c = bounds[r + 1] > 0 ? cum[bounds[r + 1] - 1] : 0ull, lo_c = bounds[r] > 0 ? cum[bounds[r] - 1] : 0ull;
        const uint32_t load = (uint32_t) (hi_c - lo_c);
        if (r == rank) summary[0] = load;
        atomicMax(&scratch[0], load);
    }
    __syncthreads();
    if (b == 0) {
        const uint32_t largest = scratch[0];
        summary[1] = largest;
        // the gate of the scatter that follows: the largest range fits a receive buffer and (if asked) the ranges
        // balance -- largest / mean <= permille / 1000
        bool ok = largest <= capacity;
        if (max_imbalance_permille != 0) ok = ok && (unsigned long long) largest * world * 1000ull <= total * max_imbalance_permille;
        summary[2] = ok ? 1u : 0u;
    }
    if (b >= 1 && b < world) summary[3 + b] = bounds[b];
}

// Barrier between the ranks of one node, stream-ordered: thread r stores `epoch` into slot `rank` of rank r's flag
// array (peer memory), then waits until slot r of the own array has reached `epoch`.  Everything this rank stored
// into peer memory in EARLIER kernels of the stream is complete before this kernel starts; the system-scope fence
// orders the flag behind it for the observer.  A peer that never arrives traps after 20 s instead of hanging the GPU.
__global__ void __launch_bounds__(EXCHANGE_MAX_RANKS)
peer_signal_wait_kernel(volatile uint32_t *flags_local, const unsigned long long *__restrict__ peer_flags, uint32_t world, uint32_t rank,
                        uint32_t epoch) {
    const uint32_t r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    uint32_t *remote = reinterpret_cast<uint32_t *>(peer_flags[r]) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    unsigned long long t_start = 0;
    for (uint32_t attempt = 0;; ++attempt) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags_local + r) : "memory");
        if ((int32_t) (v - epoch) >= 0) break;
        __nanosleep(200);
        if ((attempt & 4095u) == 4095u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t_start == 0) t_start = now;
            else if (now - t_start > 20000000000ull) __trap();
        }
    }
    __threadfence_system();
}

} // namespace vkrs
