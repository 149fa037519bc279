// vkrs_pipeline.cuh -- the pipelined one-sweep digit pass (the headline kernel).
//
// One 8-bit digit pass in a single sweep over the keys (4 B read + 4 B write per key), like
// onesweep_pass_kernel, but organised as a persistent, warp-specialised software pipeline so
// that the two latencies that stall the simple kernel -- the global load of a tile and the
// chained-scan look-back -- never sit on the workers' critical path:
//
//   control group (8 warps, one thread per digit) hands out tiles in ticket order; for tile j+1 it issues a TMA bulk
//                           copy (cp.async.bulk, completes on an mbarrier) into the other half of
//                           a double-buffered shared-memory ring, then does tile j's chained scan:
//                           publishes the tile's digit counts (AGGREGATE), looks back over earlier
//                           tiles, publishes the INCLUSIVE prefix and leaves the tile's global
//                           digit bases in shared memory.
//   worker warps            iteration j:  wait TMA(j) -> rank tile j (warp-private stable
//                           multisplit, see vkrs_tile.cuh) -> | barrier | -> digit counts of j to the
//                           control warp, tile-local scan; write tile j-1 out to global memory
//                           (its look-back had a whole ranking phase to finish) -> | barrier | ->
//                           move tile j's keys to their rank in the `sorted` staging buffer.
//
// Signals: full[2]/empty[2] (TMA ring), counts_ready[2] (workers -> control), prefix_ready[2]
// (control -> workers) are mbarriers; the workers use named barrier 1 among themselves, the 256
// digit threads named barrier 2.  Replaces multi_radixsort.comp:56-126 (every work group
// re-reading all W histograms, 3 barriers per 256 keys, 4-byte scattered stores).
#pragma once
#include "vkrs_async.cuh"
#include "vkrs_common.cuh"
#include "vkrs_kernels.cuh"

namespace vkrs {

enum TileMode : uint32_t { TILE_TMA = 0, TILE_MANUAL = 1, TILE_END = 2 };

template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT>
struct PipeGroupSmem {
    static constexpr int WARPS = WORKERS / 32;
    static constexpr int TILE = WORKERS * KPT;
    alignas(128) KeyT in[2][TILE];                       // TMA destinations (keys of tile j, j+1)
    alignas(128) uint32_t vin[HAS_VALUES ? 2 : 1][HAS_VALUES ? TILE : 4];
    alignas(128) KeyT sorted[TILE];                      // tile in digit order, staged for the write-out
    alignas(128) uint32_t sorted_v[HAS_VALUES ? TILE : 4];
    uint32_t warp_cnt[WARPS][RADIX];                     // per-warp digit counters, then exclusive bases
    uint32_t count[2][RADIX];                            // tile digit counts for the control group
    uint32_t lexcl[2][RADIX];                            // start of each digit inside the sorted tile
    uint32_t bin_dst[2][RADIX];                          // global start of the digit run minus lexcl
    uint32_t tile_id[2], tile_mode[2];
    uint32_t scan_scratch[8];
    alignas(8) uint64_t full[2], empty[2], counts_ready[2], prefix_ready[2];
};

// GROUPS worker groups share one CTA and one control group.  With GROUPS == 2 the groups run the
// same per-tile loop half a period apart ("ping-pong"): a token (named barriers 6/7) lets only one
// group at a time into the ALU-bound ranking phase, so one group's ranking overlaps the other's
// shared-memory-bound scatter / write-out instead of both fighting for the same pipe.
template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT, int GROUPS>
struct PipeSmem {
    using Group = PipeGroupSmem<KeyT, HAS_VALUES, WORKERS, KPT>;
    static constexpr int WARPS = Group::WARPS;
    static constexpr int TILE = Group::TILE;
    Group g[GROUPS];
};

constexpr int CTRL_THREADS = RADIX; // the control group: one thread per digit (8 warps)
constexpr int CTRL_WINDOW = 4;       // earlier tiles a control thread polls together

template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() {
    if constexpr (REGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() {
    if constexpr (REGS > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// REG_WORKER / REG_CTRL: per-thread register budgets after the role split (setmaxnreg; 0 = keep
// the launch allocation).  The control threads give registers back, the workers pick them up.
template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT, int MIN_BLOCKS, int REG_WORKER, int REG_CTRL,
          int MATCH = MATCH_TABLE, int GROUPS = 1>
__global__ void __launch_bounds__(GROUPS * WORKERS + CTRL_THREADS, MIN_BLOCKS)
onesweep_pipelined_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out,
                          const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out, uint32_t n,
                          uint32_t shift, const uint32_t *__restrict__ bin_start, uint32_t *status,
                          uint32_t *status_clear, uint32_t *ticket, uint32_t *error_flag,
                          unsigned long long *dbg) {
    using Smem = PipeSmem<KeyT, HAS_VALUES, WORKERS, KPT, GROUPS>;
    constexpr int WARPS = Smem::WARPS;
    constexpr uint32_t TILE = Smem::TILE;
    constexpr int ALL_WORKERS = GROUPS * WORKERS;
    static_assert(GROUPS == 1 || GROUPS == 2, "one worker group, or two in ping-pong");
    static_assert(WORKERS >= RADIX && WORKERS % 128 == 0, "whole warpgroups of workers, one thread per digit");
    static_assert(TILE <= 65536 && KPT % 2 == 0, "tile ranks are stored in 16 bits, two per register");
    extern __shared__ __align__(128) unsigned char smem_raw_pipe[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw_pipe);

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t num_tiles = (uint32_t) (((uint64_t) n + TILE - 1) / TILE);
    // TMA needs 16-byte aligned global addresses; tile offsets are multiples of 16 bytes.
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(keys_in) & 15) == 0) &&
                        (!HAS_VALUES || (reinterpret_cast<uintptr_t>(vals_in) & 15) == 0);

    if (tid == 0) {
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                mbar_init(&sm.g[gi].full[b], 1);
                mbar_init(&sm.g[gi].empty[b], WARPS);
                mbar_init(&sm.g[gi].counts_ready[b], RADIX / 32);
                mbar_init(&sm.g[gi].prefix_ready[b], CTRL_THREADS / 32);
            }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= ALL_WORKERS) {
        // ================================ control group ================================
        setmaxnreg_dec<REG_CTRL>();
        const uint32_t d = tid - ALL_WORKERS; // this thread's digit
        // Thread 0 of the group claims the next tile of worker group `s` and starts its load into
        // ring slot `slot`.
        auto claim = [&](typename Smem::Group &s, uint32_t slot) {
            const uint32_t t = atomicAdd(ticket, 1u);
            uint32_t mode = TILE_END;
            if (t < num_tiles) {
                const uint64_t base = (uint64_t) t * TILE;
                const bool full_tile = (uint64_t) n - base >= TILE;
                mode = (tma_ok && full_tile) ? TILE_TMA : TILE_MANUAL;
            }
            s.tile_id[slot] = t;
            s.tile_mode[slot] = mode;
            if (mode == TILE_TMA) {
                const uint64_t base = (uint64_t) t * TILE;
                constexpr uint32_t kbytes = TILE * sizeof(KeyT);
                mbar_arrive_expect_tx(&s.full[slot], kbytes + (HAS_VALUES ? TILE * 4u : 0u));
                bulk_copy_g2s(s.in[slot], keys_in + base, kbytes, &s.full[slot]);
                if (HAS_VALUES) bulk_copy_g2s(s.vin[slot], vals_in + base, TILE * 4u, &s.full[slot]);
            } else {
                mbar_arrive(&s.full[slot]); // the workers load (or stop) themselves
            }
        };

        // phase timers (tuning aid, only when dbg != nullptr): cycles of control thread 0
        unsigned long long t_counts = 0, t_claim = 0, t_look = 0, n_tiles = 0, t0 = 0;
        if (d == 0) {
#pragma unroll
            for (int gi = 0; gi < GROUPS; ++gi) claim(sm.g[gi], 0);
        }
        named_bar_sync(5, CTRL_THREADS);
        uint32_t mode[GROUPS], tile[GROUPS];
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi) {
            mode[gi] = sm.g[gi].tile_mode[0];
            tile[gi] = sm.g[gi].tile_id[0];
        }
        // The worker groups take turns, so their tiles reach the control group alternately too.
        for (uint32_t j = 0;; ++j) {
            bool any = false;
#pragma unroll
            for (int gi = 0; gi < GROUPS; ++gi) {
                if (mode[gi] == TILE_END) continue;
                any = true;
                typename Smem::Group &s = sm.g[gi];
                const uint32_t slot = j & 1, par = (j >> 1) & 1;
                if (dbg) t0 = clock64();
                mbar_wait_sleep(&s.counts_ready[slot], par, 128);
                if (dbg) { const unsigned long long t1 = clock64(); t_counts += t1 - t0; t0 = t1; }
                // Claim tile j+1 only now, a fixed distance (scatter of j + ranking of j+1) ahead of
                // the moment its own counts will be published: a ticket taken earlier would sit
                // unpublished for a variable time and stall every later tile's look-back.  The
                // other ring slot was drained by the scatter of tile j-1, long ago.
                if (d == 0) {
                    if (j >= 1) mbar_wait(&s.empty[slot ^ 1], ((j - 1) >> 1) & 1);
                    claim(s, slot ^ 1);
                    if (dbg) { const unsigned long long t1 = clock64(); t_claim += t1 - t0; t0 = t1; }
                }
                // ---- chained scan of tile j, this thread's digit ----
                const uint32_t cnt = s.count[slot][d];
                const ChainedScanBase base{status, status_clear, bin_start, error_flag, tile[gi]};
                base.publish(d, cnt);
                s.bin_dst[slot][d] = base.template resolve<CTRL_WINDOW, 100>(d, cnt) - s.lexcl[slot][d];
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.prefix_ready[slot]);
                if (dbg) { t_look += clock64() - t0; n_tiles++; }
                named_bar_sync(5, CTRL_THREADS); // thread 0's claim is visible to the group
                mode[gi] = s.tile_mode[slot ^ 1];
                tile[gi] = s.tile_id[slot ^ 1];
            }
            if (!any) break;
        }
        if (dbg && d == 0) {
            atomicAdd(dbg + 0, t_counts);
            atomicAdd(dbg + 1, t_claim);
            atomicAdd(dbg + 2, t_look);
            atomicAdd(dbg + 4, n_tiles);
        }
        return;
    }

    setmaxnreg_inc<REG_WORKER>();
    // ==================================== workers ====================================
    const int grp = GROUPS == 1 ? 0 : tid / WORKERS;
    const int gtid = tid - grp * WORKERS, warp = gtid >> 5;
    typename Smem::Group &s = sm.g[grp];
    const uint32_t bar_w = 1 + grp, bar_d = 3 + grp;       // this group's worker / digit barriers
    const uint32_t tok_mine = 6 + grp, tok_other = 7 - grp; // ranking tokens (GROUPS == 2)
    const uint32_t lt_mask = lanemask_lt(), gt_mask = lanemask_gt();
    const DigitBitMasks bm(sizeof(KeyT) == 8 ? (shift & 31u) : shift);
    const LaneNibbleConsts lc(lane);
    const uint32_t dsel = digit_selector(shift);
    auto key_word = [&](KeyT key) -> uint32_t { // the 32-bit word of the key that holds the digit
        return sizeof(KeyT) == 8 ? (uint32_t) ((uint64_t) key >> (shift & 32u)) : (uint32_t) key;
    };
    uint32_t *my_cnt = s.warp_cnt[warp];
    const uint32_t chunk0 = warp * (KPT * 32) + lane; // warp-striped: lane l holds chunk[i*32 + l]
    uint32_t prev_valid = 0;
    // phase timers of worker warp 0 (tuning aid): wait-for-tile, rank, barrier A, digit section,
    // wait-for-prefix, write-out, barrier B, scatter
    unsigned long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tp = 0;
    const bool timing = dbg != nullptr && tid < 32;
#define VKRS_PHASE(k)                                \
    if (timing) {                                    \
        const unsigned long long tn = clock64();     \
        tw[k] += tn - tp;                            \
        tp = tn;                                     \
    }

    auto write_out = [&](uint32_t pslot, uint32_t valid) {
        const bool full = valid == TILE;
#pragma unroll
        for (int jj = 0; jj < KPT; ++jj) {
            const uint32_t p = gtid + jj * WORKERS;
            const KeyT k = s.sorted[p];
            const uint32_t g = s.bin_dst[pslot][digit_prmt(key_word(k), dsel)] + p;
            if (full || p < valid) {
                keys_out[g] = k;
                if (HAS_VALUES) vals_out[g] = s.sorted_v[p];
            }
        }
    };

    uint32_t j = 0;
    for (;; ++j) {
        const uint32_t slot = j & 1, par = (j >> 1) & 1;
        if (timing) tp = clock64();
        mbar_wait(&s.full[slot], par);
        const uint32_t mode = s.tile_mode[slot];
        if (mode == TILE_END) break;
        const uint32_t tile = s.tile_id[slot];
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t valid = ((uint64_t) n - tile_base < TILE) ? (uint32_t) (n - tile_base) : TILE;
        const KeyT *tin = s.in[slot];
        if (mode == TILE_MANUAL) {
            // Partial last tile or a buffer TMA cannot address: the workers copy it in.  Missing
            // keys become all-ones: digit 255 at every shift and last in memory order, so they
            // rank after every real key, at tile positions >= valid.
            for (uint32_t p = gtid; p < TILE; p += WORKERS) {
                s.in[slot][p] = p < valid ? ld_stream(keys_in + tile_base + p) : ~KeyT(0);
                if (HAS_VALUES) s.vin[slot][p] = p < valid ? ld_stream(vals_in + tile_base + p) : 0u;
            }
            named_bar_sync(bar_w, WORKERS);
        }
        // ---- ping-pong: wait for the other group to leave the ranking phase ----
        if (GROUPS == 2 && (grp == 1 || j > 0)) named_bar_sync(tok_mine, ALL_WORKERS);
        VKRS_PHASE(0)

        // ---- rank inside the warp (see vkrs_tile.cuh for the protocol) ----
#pragma unroll
        for (int q = 0; q < RADIX / 32; ++q) my_cnt[lane + 32 * q] = 0;
        __syncwarp();
        uint32_t rank2[KPT / 2];
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const uint32_t word = key_word(tin[chunk0 + i * 32]);
            const uint32_t d = digit_prmt(word, dsel);
            const uint32_t peers = MATCH == MATCH_TABLE ? match_key_table(word, d, bm, lc) : match_key_ptx(word, bm);
            const uint32_t r = my_cnt[d] + __popc(peers & lt_mask);
            if ((peers & gt_mask) == 0) my_cnt[d] = r + 1; // highest lane of the group
            if (i & 1) rank2[i / 2] |= r << 16;
            else rank2[i / 2] = r;
            __syncwarp();
        }
        if (GROUPS == 2) named_bar_arrive(tok_other, ALL_WORKERS); // the other group may rank now
        VKRS_PHASE(1)
        named_bar_sync(bar_w, WORKERS); // (A) all warp counters final; sorted[] holds tile j-1 completely
        VKRS_PHASE(2)

        // ---- digit threads: tile counts to the control group, tile-local scan, warp bases ----
        if (gtid < RADIX) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) total += s.warp_cnt[w][gtid];
            // block-wide exclusive scan over the 256 digit threads
            const uint32_t incl = warp_inclusive_scan(total, lane);
            if (lane == 31) s.scan_scratch[warp] = incl;
            named_bar_sync(bar_d, RADIX);
            uint32_t warp_prefix = 0;
#pragma unroll
            for (int w = 0; w < RADIX / 32; ++w)
                if (w < warp) warp_prefix += s.scan_scratch[w];
            const uint32_t local_excl = warp_prefix + incl - total;
            uint32_t running = local_excl;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = s.warp_cnt[w][gtid];
                s.warp_cnt[w][gtid] = running;
                running += c;
            }
            // what the rest of the grid must see: real keys only (padding sits in digit 255)
            s.count[slot][gtid] = (valid != TILE && gtid == RADIX - 1) ? total - (TILE - valid) : total;
            s.lexcl[slot][gtid] = local_excl;
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.counts_ready[slot]);
        }

        VKRS_PHASE(3)
        // ---- write tile j-1 out: its look-back ran while we ranked tile j ----
        if (j > 0) {
            mbar_wait(&s.prefix_ready[slot ^ 1], ((j - 1) >> 1) & 1);
            VKRS_PHASE(4)
            write_out(slot ^ 1, prev_valid);
        }
        VKRS_PHASE(5)
        named_bar_sync(bar_w, WORKERS); // (B) warp bases of tile j ready; sorted[] free
        VKRS_PHASE(6)

        // ---- keys (and payloads) of tile j to their rank in the staging buffer ----
        // In batches, all loads of a batch before its stores: the three shared-memory accesses
        // per key are latency-bound if they run key by key.
        constexpr int SB = KPT % 8 == 0 ? 8 : (KPT % 4 == 0 ? 4 : 2);
#pragma unroll
        for (int i0 = 0; i0 < KPT; i0 += SB) {
            KeyT kb[SB];
            uint32_t rb[SB];
#pragma unroll
            for (int i = 0; i < SB; ++i) kb[i] = tin[chunk0 + (i0 + i) * 32];
#pragma unroll
            for (int i = 0; i < SB; ++i) rb[i] = my_cnt[digit_prmt(key_word(kb[i]), dsel)];
#pragma unroll
            for (int i = 0; i < SB; ++i) {
                const int k = i0 + i;
                rb[i] += (k & 1) ? (rank2[k / 2] >> 16) : (rank2[k / 2] & 0xffffu);
            }
#pragma unroll
            for (int i = 0; i < SB; ++i) s.sorted[rb[i]] = kb[i];
            if (HAS_VALUES) {
                uint32_t vb[SB];
#pragma unroll
                for (int i = 0; i < SB; ++i) vb[i] = s.vin[slot][chunk0 + (i0 + i) * 32];
#pragma unroll
                for (int i = 0; i < SB; ++i) s.sorted_v[rb[i]] = vb[i];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[slot]); // ring slot may be refilled
        VKRS_PHASE(7)
        prev_valid = valid;
    }
    // Out of tiles.  The other group may still hold one claimed tile: hand it the ranking token a
    // last time (see the header comment; pending arrivals on a named barrier are harmless at exit).
    if (GROUPS == 2) named_bar_arrive(tok_other, ALL_WORKERS);
    if (timing && lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(dbg + 8 + k, tw[k]);
    }
#undef VKRS_PHASE
    if (j > 0) { // drain: the last tile this group ranked
        named_bar_sync(bar_w, WORKERS);
        mbar_wait(&s.prefix_ready[(j - 1) & 1], ((j - 1) >> 1) & 1);
        write_out((j - 1) & 1, prev_valid);
    }
}

} // namespace vkrs
