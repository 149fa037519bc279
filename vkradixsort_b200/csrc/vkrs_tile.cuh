// vkrs_tile.cuh -- the per-tile stable multisplit that every scatter kernel is built on.
//
// One CTA takes a tile of TILE = THREADS*KPT consecutive keys and stably partitions it by
// the current 8-bit digit:
//   1. each warp owns a contiguous chunk of KPT*32 keys, loaded warp-striped (lane l holds
//      chunk[i*32 + l]) so every load instruction is one coalesced 128-byte line and the
//      memory order inside the chunk is (round i, lane l);
//   2. per round, lanes with the same digit find each other (match_digit), the lowest lane of
//      each group bumps the warp's private digit counter in shared memory and broadcasts the
//      old value: rank-in-warp = old + #lower lanes of the group  (stable by construction;
//      this replaces the reference's bin_flags bit matrix, multi_radixsort.comp:97-118);
//   3. one thread per digit turns the WARPS x 256 counters into exclusive prefixes over warps,
//      and a block scan over digits gives each digit's start inside the sorted tile;
//   4. the caller-supplied BaseFn maps (digit, tile count) -> global start of this tile's
//      run for that digit (chained look-back in the fused path, running shared offsets in the
//      staged / single paths) -- the role of global_offsets[] in multi_radixsort.comp:75-76,120-122;
//   5. keys go to their rank in a shared-memory copy of the tile, then are written out in
//      tile order, so a warp writes long runs of consecutive addresses per digit instead of
//      the reference's one-key-per-sector scatter (multi_radixsort.comp:119).
#pragma once
#include "vkrs_common.cuh"

namespace vkrs {

template <typename KeyT, int THREADS, int KPT>
struct TileSmem {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * KPT;
    uint32_t warp_cnt[WARPS][RADIX]; // per-warp digit counters, later exclusive bases
    KeyT tile[TILE];                 // keys (then payloads) in tile-sorted order
    uint32_t bin_dst[RADIX];         // global start of the digit run minus its start inside the tile
    uint32_t scan_scratch[8];
    uint32_t misc[4];
};

template <typename KeyT, bool HAS_VALUES, int THREADS, int KPT, int MATCH>
struct TileSorter {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * KPT;
    static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per digit is required");
    static_assert(TILE <= 65536, "tile ranks are stored in 16 bits");
    using Smem = TileSmem<KeyT, THREADS, KPT>;

    // keys_in/vals_in point at the first key of the tile; `valid` (<= TILE) keys exist.
    // `base` is called by threads 0..255 (digit == threadIdx.x):
    //   base.publish(digit, count)  as soon as the tile's digit counts are known (the fused path
    //                               posts its AGGREGATE status word here, so later tiles can move on);
    //   base.resolve(digit, count)  as late as possible -- after this thread's keys are already in
    //                               their tile slots -- returns the global index where this tile's
    //                               run of `digit` starts (the fused path's look-back; by now the
    //                               earlier tiles have had time to publish, so the wait is short).
    template <class Base>
    static __device__ __forceinline__ void run(Smem &s, const KeyT *keys_in, KeyT *keys_out, const uint32_t *vals_in,
                                               uint32_t *vals_out, uint32_t valid, uint32_t shift, Base base) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const bool full = valid == TILE;

        // ---- load (warp-striped) ----
        KeyT key[KPT];
        uint32_t val[HAS_VALUES ? KPT : 1];
        const uint32_t chunk0 = warp * (KPT * 32) + lane;
        if (full) {
#pragma unroll
            for (int i = 0; i < KPT; ++i) key[i] = ld_stream(keys_in + chunk0 + i * 32);
            if (HAS_VALUES) {
#pragma unroll
                for (int i = 0; i < KPT; ++i) val[i] = ld_stream(vals_in + chunk0 + i * 32);
            }
        } else {
            // Missing keys become all-ones: digit 255 at every shift and, being last in memory
            // order, they rank after every real key, i.e. at tile positions >= valid.
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const uint32_t idx = chunk0 + i * 32;
                key[i] = idx < valid ? ld_stream(keys_in + idx) : ~KeyT(0);
                if (HAS_VALUES) val[i] = idx < valid ? ld_stream(vals_in + idx) : 0u;
            }
        }

        // ---- zero this warp's counters ----
#pragma unroll
        for (int j = 0; j < RADIX / 32; ++j) s.warp_cnt[warp][lane + 32 * j] = 0;
        __syncwarp();

        // ---- rank inside the warp ----
        // Per round every lane reads the warp's running count of its digit, adds the number of
        // lower lanes holding the same digit, and the highest lane of each group stores the
        // bumped count back -- the role of `prefix == count - 1` in multi_radixsort.comp:120-122.
        // ranks are < TILE <= 65536: two per register
        static_assert(KPT % 2 == 0, "ranks are packed in pairs");
        uint32_t rank2[KPT / 2];
        const uint32_t lt_mask = lanemask_lt(), gt_mask = lanemask_gt();
        const DigitBitMasks bm(sizeof(KeyT) == 8 ? (shift & 31u) : shift);
        uint32_t *my_cnt = s.warp_cnt[warp];
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const uint32_t d = digit_of(key[i], shift);
            uint32_t peers;
            if (MATCH == MATCH_PTX) {
                const uint32_t word = sizeof(KeyT) == 8 ? (uint32_t) ((uint64_t) key[i] >> (shift & 32u)) : (uint32_t) key[i];
                peers = match_key_ptx(word, bm);
            } else {
                peers = match_digit<MATCH>(d);
            }
            const uint32_t r = my_cnt[d] + __popc(peers & lt_mask);
            if ((peers & gt_mask) == 0) my_cnt[d] = r + 1; // highest lane of the group
            if (i & 1) rank2[i / 2] |= r << 16;
            else rank2[i / 2] = r;
            __syncwarp(); // counter update visible to the next round
        }
        __syncthreads();

        // ---- digit threads: counters -> exclusive bases, tile counts, global bases ----
        uint32_t total = 0;
        if (tid < RADIX) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) total += s.warp_cnt[w][tid];
        }
        uint32_t counted = total; // what the rest of the grid must see: real keys only
        if (!full && tid == RADIX - 1) counted -= (TILE - valid);
        if (tid < RADIX) base.publish((uint32_t) tid, counted);
        const uint32_t local_excl = block_exclusive_scan_256(tid < RADIX ? total : 0u, s.scan_scratch, nullptr);
        if (tid < RADIX) {
            uint32_t running = local_excl;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { // re-read instead of keeping WARPS counts live in registers
                const uint32_t c = s.warp_cnt[w][tid];
                s.warp_cnt[w][tid] = running;
                running += c;
            }
        }
        __syncthreads();

        // ---- keys to their tile rank ----
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const uint32_t d = digit_of(key[i], shift);
            const uint32_t r = ((i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xffffu)) + s.warp_cnt[warp][d];
            if (HAS_VALUES) { // keep the final tile position for the payload
                if (i & 1) rank2[i / 2] = (rank2[i / 2] & 0xffffu) | (r << 16);
                else rank2[i / 2] = (rank2[i / 2] & 0xffff0000u) | r;
            }
            s.tile[r] = key[i];
        }
        if (tid < RADIX) s.bin_dst[tid] = base.resolve((uint32_t) tid, counted) - local_excl;
        __syncthreads();

        // ---- write out in tile order ----
        uint32_t dst[HAS_VALUES ? KPT : 1];
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const uint32_t p = tid + j * THREADS;
            const KeyT k = s.tile[p];
            const uint32_t g = s.bin_dst[digit_of(k, shift)] + p;
            if (HAS_VALUES) dst[j] = g;
            if (full || p < valid) keys_out[g] = k;
        }
        if (HAS_VALUES) {
            __syncthreads();
            uint32_t *vtile = reinterpret_cast<uint32_t *>(s.tile);
#pragma unroll
            for (int i = 0; i < KPT; ++i) vtile[(i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xffffu)] = val[i];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                const uint32_t p = tid + j * THREADS;
                if (full || p < valid) vals_out[dst[j]] = vtile[p];
            }
        }
    }
};

} // namespace vkrs
