// vkrs_segmented.cuh -- the segmented two-kernel digit pass: the reference's own decomposition
// (multi_radixsort_histograms.comp + multi_radixsort.comp) re-tiled for B200.
//
//   segment_histogram_kernel   CTA g counts the digits of segment g (a contiguous slab of tiles,
//                              exactly the reference's "work group w owns nb*256 consecutive keys")
//                              into row g of the histogram matrix.  One coalesced read of the keys.
//   segmented_scatter_kernel   persistent CTA, two worker groups, group g owns segment g.  Prologue:
//                              offsets[g][d] = sum_{d'<d} sum_{g'} hist[g'][d'] + sum_{g'<g} hist[g'][d]
//                              (multi_radixsort.comp:56-77) with G = 2 x #SMs rows instead of the
//                              reference's W = N/(nb*256).  Then the group walks its segment tile by
//                              tile with running per-digit offsets (global_offsets[], :120-122): TMA
//                              double-buffered loads, warp-private stable ranking, a shared-memory
//                              reorder and coalesced write-out.
//
// Why not the single-sweep chained scan here: on B200 a tile retires every ~20 ns across the chip
// while one L2 round trip takes ~500 ns, so every tile's look-back has to cover dozens of
// predecessors; the chain, not the data path, sets the pace (measured: 0.43 ms/pass however the
// tiles are shaped).  Counting first costs one extra 4 B/key read per pass and removes every
// inter-CTA dependency from the scatter.  A CTA hosts GROUPS independent worker groups (own segment,
// own ring, own named barriers): while one group is in the ALU-bound ranking phase another is in its
// shared-memory-bound scatter / write-out, so the phases of different groups overlap on the SM.
// (An explicit ping-pong token between two groups was measured slower than letting them drift.)
#pragma once
#include "vkrs_async.cuh"
#include "vkrs_common.cuh"

namespace vkrs {

// Which 8-bit "digit" a pass distributes by.
//   PARTITION == false : byte (shift/8) of the key, the reference's (key >> g_shift) & 255
//                        (multi_radixsort.comp:100); one PRMT.
//   PARTITION == true  : min(255, (key - base) >> shift) for an arbitrary shift -- the bucket index
//                        of the multi-GPU exchange (vkrs_partition): an order-preserving map of the
//                        occupied key range onto 256 buckets.  32-bit keys only.
template <typename KeyT, bool PARTITION>
struct DigitOf {
    uint32_t shift, base, dsel;
    __device__ __forceinline__ DigitOf(uint32_t shift_, uint32_t base_) : shift(shift_), base(base_), dsel(digit_selector(shift_)) {}
    // digit of the key
    __device__ __forceinline__ uint32_t operator()(KeyT key) const {
        if (PARTITION) {
            const uint32_t f = ((uint32_t) key - base) >> shift;
            return f < 255u ? f : 255u;
        }
        return digit_prmt(sizeof(KeyT) == 8 ? (uint32_t) ((uint64_t) key >> (shift & 32u)) : (uint32_t) key, dsel);
    }
    // the word the ballots test (bit masks from bit_masks()) for a key whose digit is d
    __device__ __forceinline__ uint32_t match_word(KeyT key, uint32_t d) const {
        if (PARTITION) return d;
        return sizeof(KeyT) == 8 ? (uint32_t) ((uint64_t) key >> (shift & 32u)) : (uint32_t) key;
    }
    __device__ __forceinline__ DigitBitMasks bit_masks() const { return DigitBitMasks(PARTITION ? 0u : (shift & 31u)); }
};

// Tiles [first, first + count) of segment g when `num_tiles` tiles are dealt to `num_segments`
// segments as evenly as possible (host and device agree through this one function).
__host__ __device__ __forceinline__ void segment_tiles(uint32_t g, uint32_t num_segments, uint32_t num_tiles,
                                                       uint32_t &first, uint32_t &count) {
    const uint32_t q = num_tiles / num_segments, r = num_tiles % num_segments;
    first = g * q + (g < r ? g : r);
    count = q + (g < r ? 1u : 0u);
}

// =====================================================================================
// Kernel 1: hist[g][d] = #{keys of segment g with digit d}.  Grid = number of segments.
// Shared-memory counters are lane-private 32-bit columns (bank == lane: one wavefront per atomic
// whatever the key distribution, all-equal keys included).
// =====================================================================================
constexpr int SEGHIST_THREADS = 512;

template <typename KeyT, bool PARTITION, int XF_IN = 0>
__global__ void __launch_bounds__(SEGHIST_THREADS)
segment_histogram_kernel(const KeyT *__restrict__ keys, uint32_t n, uint32_t shift, uint32_t key_base,
                         uint32_t tile_keys, uint32_t num_tiles, uint32_t *__restrict__ hist,
                         const uint32_t *__restrict__ gate) {
    __shared__ uint32_t cnt[RADIX * 32]; // [digit][lane]: bank == lane, one wavefront per atomic
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t first, count;
    segment_tiles(blockIdx.x, gridDim.x, num_tiles, first, count);
    const uint64_t lo = (uint64_t) first * tile_keys;
    uint64_t hi = lo + (uint64_t) count * tile_keys;
    if (hi > n) hi = n;
    const DigitOf<KeyT, PARTITION> digit(shift, key_base);
    uint32_t *my_col = cnt + lane;
    auto count_key = [&](KeyT k) { atomicAdd(my_col + digit(KeyXform<KeyT, XF_IN>::fwd(k)) * 32, 1u); };
    for (int i = tid; i < RADIX * 32; i += SEGHIST_THREADS) cnt[i] = 0;
    grid_dependency_wait(); // programmatic dependent launch: everything above overlaps the previous kernel's tail
    if (gate && *gate == 0) return; // conditional pass (fallback of the bucket schedule, vkrs_msd.cuh): not needed
    __syncthreads();
    if (lo < hi) {
        constexpr int VEC = 16 / sizeof(KeyT);
        const KeyT *base = keys + lo;
        const uint64_t cnt_keys = hi - lo;
        uint64_t head = ((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15) / sizeof(KeyT);
        if (head > cnt_keys) head = cnt_keys;
        const uint64_t nvec = (cnt_keys - head) / VEC;
        if (tid < head) count_key(base[tid]);
        const uint4 *vbase = reinterpret_cast<const uint4 *>(base + head);
        uint64_t v = tid;
        for (; v + 3 * SEGHIST_THREADS < nvec; v += 4 * SEGHIST_THREADS) { // four 128-bit loads in flight
            uint4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = ld_stream(vbase + v + u * SEGHIST_THREADS);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (sizeof(KeyT) == 4) {
                    count_key((KeyT) a[u].x); count_key((KeyT) a[u].y); count_key((KeyT) a[u].z); count_key((KeyT) a[u].w);
                } else {
                    count_key((KeyT) (((uint64_t) a[u].y << 32) | a[u].x));
                    count_key((KeyT) (((uint64_t) a[u].w << 32) | a[u].z));
                }
            }
        }
        for (; v < nvec; v += SEGHIST_THREADS) {
            const uint4 a = ld_stream(vbase + v);
            if (sizeof(KeyT) == 4) {
                count_key((KeyT) a.x); count_key((KeyT) a.y); count_key((KeyT) a.z); count_key((KeyT) a.w);
            } else {
                count_key((KeyT) (((uint64_t) a.y << 32) | a.x)); count_key((KeyT) (((uint64_t) a.w << 32) | a.z));
            }
        }
        const uint64_t tail0 = head + nvec * VEC;
        if (tail0 + tid < cnt_keys) count_key(base[tail0 + tid]);
    }
    __syncthreads();
    if (tid < RADIX) {
        const uint32_t *row = &cnt[tid * 32];
        uint32_t total = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) total += row[(j + tid) & 31]; // skewed: conflict-free
        hist[(size_t) blockIdx.x * RADIX + tid] = total;
    }
}

// =====================================================================================
// Kernel 2: the scatter.  GROUPS * WORKERS worker threads + one producer warp.
// =====================================================================================
template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT>
struct SegGroupSmem {
    static constexpr int WARPS = WORKERS / 32;
    static constexpr int TILE = WORKERS * KPT;
    alignas(128) KeyT in[2][TILE];                       // TMA destinations (tile j, j+1 of the segment)
    alignas(128) KeyT sorted[TILE];                      // tile in digit order, staged for the write-out
    alignas(128) uint32_t sorted_v[HAS_VALUES ? TILE : 4];
    uint32_t warp_cnt[WARPS][RADIX];                     // per-warp digit counters, then exclusive bases
    uint32_t bin_dst[2][RADIX];                          // global start of the digit run minus its start in the tile
    uint32_t scan_scratch[8];
    alignas(8) uint64_t full[2], empty[2];
    alignas(8) unsigned long long dst_ptr[2][RADIX];     // P2P mode: per-bucket destination arrays (keys, payloads)
    uint32_t lex[RADIX];                                 // P2P mode: tile-local start of every bucket
};

template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT, int GROUPS>
struct SegSmem {
    using Group = SegGroupSmem<KeyT, HAS_VALUES, WORKERS, KPT>;
    Group g[GROUPS];
};

// P2P == true (multi-GPU exchange fused into the partition): there is no single output array.
// Contiguous ranges of buckets belong to one owner rank.  dst_tables (4 x 256 x uint64, indexed by
// bucket b): [b] / [256+b] = address where THIS rank's part starts inside the key / payload receive
// buffer of b's owner -- a peer GPU's memory mapped over NVLink (or this GPU's own); [512+b] / [768+b] =
// first bucket / one past the last bucket of b's owner.  The tile is ranked and staged exactly as in
// the local case; in the staged tile the keys of one owner are contiguous, and they are written out
// as ONE run per owner and tile (tile-major layout inside the sender's part of the receive buffer:
// the receiver sorts anyway, and equal keys keep their order).  Long runs are what NVLink stores
// need; the exchange costs no extra pass and overlaps the ranking of other tiles.
// XF_IN / XF_OUT: key transform applied to every key as it is read (first pass of a typed sort) /
// undone as it is written (last pass); see KeyXform.
template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT, int GROUPS, int MIN_BLOCKS, bool PARTITION = false,
          bool P2P = false, int XF_IN = 0, int XF_OUT = 0>
__global__ void __launch_bounds__(GROUPS * WORKERS + 32, MIN_BLOCKS)
segmented_scatter_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out,
                         const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out, uint32_t n,
                         uint32_t shift, uint32_t key_base, const uint32_t *__restrict__ hist, uint32_t num_tiles,
                         unsigned long long *dbg, const unsigned long long *__restrict__ dst_tables,
                         const uint32_t *__restrict__ gate) {
    using Smem = SegSmem<KeyT, HAS_VALUES, WORKERS, KPT, GROUPS>;
    using Group = typename Smem::Group;
    constexpr int WARPS = Group::WARPS;
    constexpr uint32_t TILE = Group::TILE;
    constexpr int ALL_WORKERS = GROUPS * WORKERS;
    static_assert(GROUPS >= 1 && GROUPS <= 4, "1..4 worker groups per CTA");
    static_assert(WORKERS >= RADIX && WORKERS % 32 == 0, "one worker thread per digit is required");
    static_assert(TILE <= 65536 && KPT % 2 == 0, "tile ranks are stored in 16 bits, two per register");
    extern __shared__ __align__(128) unsigned char smem_raw_seg[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw_seg);

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t num_segments = gridDim.x * GROUPS;
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(keys_in) & 15) == 0) &&
                        (!HAS_VALUES || (reinterpret_cast<uintptr_t>(vals_in) & 15) == 0);

    if (tid == 0) {
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                mbar_init(&sm.g[gi].full[b], 1);
                mbar_init(&sm.g[gi].empty[b], WARPS);
            }
        mbar_fence_init();
    }
    grid_dependency_wait(); // the histogram matrix and (in later passes) the keys come from earlier kernels
    __syncthreads();
    if (gate && *gate == 0) return; // conditional pass (fallback of the bucket schedule, vkrs_msd.cuh): not needed

    // A tile can go through TMA when it is full and its global address is 16-byte aligned.
    auto tile_is_tma = [&](uint32_t tile) { return tma_ok && (uint64_t) n - (uint64_t) tile * TILE >= TILE; };

    if (tid >= ALL_WORKERS) {
        // ============================ producer warp (lane 0) ============================
        if (lane != 0) return;
        uint32_t first[GROUPS], count[GROUPS];
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi) segment_tiles(blockIdx.x * GROUPS + gi, num_segments, num_tiles, first[gi], count[gi]);
        uint32_t max_count = 0;
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi) max_count = count[gi] > max_count ? count[gi] : max_count;
        for (uint32_t j = 0; j < max_count; ++j) {
#pragma unroll
            for (int gi = 0; gi < GROUPS; ++gi) {
                if (j >= count[gi]) continue;
                Group &s = sm.g[gi];
                const uint32_t slot = j & 1, tile = first[gi] + j;
                if (j >= 2) mbar_wait_sleep(&s.empty[slot], ((j - 2) >> 1) & 1, 64); // tile j-2 fully consumed
                if (tile_is_tma(tile)) {
                    const uint64_t base = (uint64_t) tile * TILE;
                    constexpr uint32_t kbytes = TILE * sizeof(KeyT);
                    mbar_arrive_expect_tx(&s.full[slot], kbytes);
                    bulk_copy_g2s(s.in[slot], keys_in + base, kbytes, &s.full[slot]);
                } else {
                    mbar_arrive(&s.full[slot]); // the workers copy this tile in themselves
                }
            }
        }
        return;
    }

    // ==================================== workers ====================================
    const int grp = GROUPS == 1 ? 0 : tid / WORKERS;
    const int gtid = tid - grp * WORKERS, warp = gtid >> 5;
    Group &s = sm.g[grp];
    const uint32_t seg = blockIdx.x * GROUPS + grp;
    uint32_t first_tile, tile_count;
    segment_tiles(seg, num_segments, num_tiles, first_tile, tile_count);
    const uint32_t bar_w = 1 + grp, bar_d = 1 + GROUPS + grp; // this group's worker / digit named barriers
    const uint32_t lt_mask = lanemask_lt(), gt_mask = lanemask_gt();
    const DigitOf<KeyT, PARTITION> digit(shift, key_base);
    const DigitBitMasks bm = digit.bit_masks();
    const LaneNibbleConsts lc(lane);
    uint32_t *my_cnt = s.warp_cnt[warp];
    const uint32_t chunk0 = warp * (KPT * 32) + lane; // warp-striped: lane l holds chunk[i*32 + l]
    // The digit threads are the group's first eight warps (moving them to the last eight, which the
    // scheduler favours, was measured 4 % slower).
    const bool is_digit_thread = gtid < RADIX;
    const uint32_t dgt = gtid, dwarp = dgt >> 5; // this thread's digit, its warp among the eight

    // ---- prologue (multi_radixsort.comp:56-77): this segment's first output index per digit ----
    uint32_t running_base = 0; // digit thread `gtid`: where the next tile's run of that digit starts
    uint32_t owner_first = 0, owner_end = 0; // P2P: the bucket range of this digit's owner rank
    if (is_digit_thread) {
        uint32_t below = 0, total = 0;
        uint32_t g2 = 0;
        for (; g2 + 8 <= num_segments; g2 += 8) { // eight independent loads in flight
            uint32_t h[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) h[u] = __ldcg(hist + (size_t) (g2 + u) * RADIX + dgt);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                total += h[u];
                if (g2 + u < seg) below += h[u];
            }
        }
        for (; g2 < num_segments; ++g2) {
            const uint32_t h = __ldcg(hist + (size_t) g2 * RADIX + dgt);
            total += h;
            if (g2 < seg) below += h;
        }
        const uint32_t incl = warp_inclusive_scan(total, lane);
        if (lane == 31) s.scan_scratch[dwarp] = incl;
        named_bar_sync(bar_d, RADIX);
        uint32_t warp_prefix = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; ++w)
            if (w < (int) dwarp) warp_prefix += s.scan_scratch[w];
        // local output: digits are laid out one after the other
        running_base = warp_prefix + incl - total + below;
        if (P2P) {
            // one running offset per OWNER (kept redundantly by all of its digit threads): where this
            // segment starts inside the sender's part = keys of earlier segments going to that owner
            s.dst_ptr[0][dgt] = dst_tables[dgt];
            if (HAS_VALUES) s.dst_ptr[1][dgt] = dst_tables[RADIX + dgt];
            owner_first = (uint32_t) dst_tables[2 * RADIX + dgt];
            owner_end = (uint32_t) dst_tables[3 * RADIX + dgt];
            s.lex[dgt] = below;
            named_bar_sync(bar_d, RADIX);
            running_base = 0;
            for (uint32_t b = owner_first; b < owner_end; ++b) running_base += s.lex[b];
        }
        named_bar_sync(bar_d, RADIX); // scan_scratch / lex are reused by the per-tile scan
    }

    // phase timers of worker warp 0 (tuning aid): wait-for-tile(+token), rank, barrier A, digit
    // section, -, write-out, barrier B, scatter
    unsigned long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tp = 0;
    const bool timing = dbg != nullptr && tid < 32;
#define VKRS_PHASE(k)                                \
    if (timing) {                                    \
        const unsigned long long tn = clock64();     \
        tw[k] += tn - tp;                            \
        tp = tn;                                     \
    }

    auto write_one = [&](uint32_t pslot, uint32_t p) {
        const KeyT k = s.sorted[p];
        const uint32_t d = digit(k);
        const uint32_t g = s.bin_dst[pslot][d] + p;
        if (P2P) {
            reinterpret_cast<KeyT *>(s.dst_ptr[0][d])[g] = KeyXform<KeyT, XF_OUT>::inv(k);
            if (HAS_VALUES) reinterpret_cast<uint32_t *>(s.dst_ptr[1][d])[g] = s.sorted_v[p];
        } else {
            keys_out[g] = KeyXform<KeyT, XF_OUT>::inv(k);
            if (HAS_VALUES) vals_out[g] = s.sorted_v[p];
        }
    };
    auto write_out = [&](uint32_t pslot, uint32_t valid) {
        if (valid == TILE) { // the usual case, without a bounds check (= a branch) per key
#pragma unroll
            for (int jj = 0; jj < KPT; ++jj) write_one(pslot, gtid + jj * WORKERS);
        } else {
#pragma unroll
            for (int jj = 0; jj < KPT; ++jj) {
                const uint32_t p = gtid + jj * WORKERS;
                if (p < valid) write_one(pslot, p);
            }
        }
    };

    uint32_t prev_valid = 0;
    for (uint32_t j = 0; j < tile_count; ++j) {
        const uint32_t slot = j & 1, par = (j >> 1) & 1;
        const uint32_t tile = first_tile + j;
        const uint64_t tile_base = (uint64_t) tile * TILE;
        const uint32_t valid = ((uint64_t) n - tile_base < TILE) ? (uint32_t) (n - tile_base) : TILE;
        const KeyT *tin = s.in[slot];
        if (timing) tp = clock64();
        mbar_wait(&s.full[slot], par);
        if (!tile_is_tma(tile)) {
            // Partial last tile or a buffer TMA cannot address: the workers copy it in.  Missing
            // keys become all-ones: digit 255 at every shift and last in memory order, so they
            // rank after every real key, at tile positions >= valid.
            for (uint32_t p = gtid; p < TILE; p += WORKERS) {
                s.in[slot][p] = p < valid ? ld_stream(keys_in + tile_base + p) : KeyXform<KeyT, XF_IN>::inv(~KeyT(0));
            }
            named_bar_sync(bar_w, WORKERS);
        }
        VKRS_PHASE(0)

        // ---- rank inside the warp (protocol: see vkrs_tile.cuh) ----
#pragma unroll
        for (int q = 0; q < RADIX / 32; ++q) my_cnt[lane + 32 * q] = 0;
        __syncwarp();
        uint32_t rank2[KPT / 2];
        // Software-pipelined: the ballots of round i+1 are issued before the counter update of
        // round i, so the shared-memory round trip of one round hides behind the votes of the next.
        const KeyT key0 = KeyXform<KeyT, XF_IN>::fwd(tin[chunk0]);
        uint32_t d_cur = digit(key0);
        uint32_t peers_cur = match_key_table(digit.match_word(key0, d_cur), d_cur, bm, lc);
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            uint32_t d_next = 0, peers_next = 0;
            if (i + 1 < KPT) {
                const KeyT key = KeyXform<KeyT, XF_IN>::fwd(tin[chunk0 + (i + 1) * 32]);
                d_next = digit(key);
                peers_next = match_key_table(digit.match_word(key, d_next), d_next, bm, lc);
            }
            const uint32_t r = my_cnt[d_cur] + __popc(peers_cur & lt_mask);
            if ((peers_cur & gt_mask) == 0) my_cnt[d_cur] = r + 1; // highest lane of the group
            if (i & 1) rank2[i / 2] |= r << 16;
            else rank2[i / 2] = r;
            __syncwarp();
            d_cur = d_next;
            peers_cur = peers_next;
        }
        // Payloads do not go through the shared-memory ring: they are needed once, at the scatter, so
        // each thread fetches its own (coalesced, warp-striped like the keys) straight into registers
        // now and the loads complete under the digit section and the write-out of tile j-1.
        uint32_t vreg[HAS_VALUES ? KPT : 1];
        if (HAS_VALUES) {
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const uint32_t idx = chunk0 + i * 32;
                vreg[i] = (valid == TILE || idx < valid) ? ld_stream(vals_in + tile_base + idx) : 0u;
            }
        }
        VKRS_PHASE(1)
        named_bar_sync(bar_w, WORKERS); // (A) all warp counters final; sorted[] holds tile j-1 completely
        VKRS_PHASE(2)

        // ---- digit threads: tile-local scan, warp bases, this tile's global digit bases ----
        if (is_digit_thread) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) total += s.warp_cnt[w][dgt];
            const uint32_t incl = warp_inclusive_scan(total, lane);
            if (lane == 31) s.scan_scratch[dwarp] = incl;
            named_bar_sync(bar_d, RADIX);
            uint32_t warp_prefix = 0;
#pragma unroll
            for (int w = 0; w < RADIX / 32; ++w)
                if (w < (int) dwarp) warp_prefix += s.scan_scratch[w];
            const uint32_t local_excl = warp_prefix + incl - total;
            uint32_t running = local_excl;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = s.warp_cnt[w][dgt];
                s.warp_cnt[w][dgt] = running;
                running += c;
            }
            if (P2P) {
                // the owner's keys form one chunk [start, end) of the staged tile -> one run in its buffer
                s.lex[dgt] = local_excl;
                named_bar_sync(bar_d, RADIX);
                const uint32_t start = s.lex[owner_first];
                const uint32_t end = owner_end == RADIX ? valid : s.lex[owner_end]; // padding sits after the real keys
                s.bin_dst[slot][dgt] = running_base - start;
                running_base += end - start;
            } else {
                s.bin_dst[slot][dgt] = running_base - local_excl;
                // real keys only: the padding of a partial tile sits in digit 255
                running_base += (valid != TILE && dgt == RADIX - 1) ? total - (TILE - valid) : total;
            }
        }
        VKRS_PHASE(3)
        // ---- write tile j-1 out (overlaps the digit threads' work above) ----
        if (j > 0) write_out(slot ^ 1, prev_valid);
        VKRS_PHASE(5)
        named_bar_sync(bar_w, WORKERS); // (B) warp bases and bin_dst of tile j ready; sorted[] free
        VKRS_PHASE(6)

        // ---- keys (and payloads) of tile j to their rank in the staging buffer ----
        constexpr int SB = KPT % 8 == 0 ? 8 : (KPT % 4 == 0 ? 4 : 2);
#pragma unroll
        for (int i0 = 0; i0 < KPT; i0 += SB) {
            KeyT kb[SB];
            uint32_t rb[SB];
#pragma unroll
            for (int i = 0; i < SB; ++i) kb[i] = KeyXform<KeyT, XF_IN>::fwd(tin[chunk0 + (i0 + i) * 32]);
#pragma unroll
            for (int i = 0; i < SB; ++i) rb[i] = my_cnt[digit(kb[i])];
#pragma unroll
            for (int i = 0; i < SB; ++i) {
                const int k = i0 + i;
                rb[i] += (k & 1) ? (rank2[k / 2] >> 16) : (rank2[k / 2] & 0xffffu);
            }
#pragma unroll
            for (int i = 0; i < SB; ++i) s.sorted[rb[i]] = kb[i];
            if (HAS_VALUES) {
#pragma unroll
                for (int i = 0; i < SB; ++i) s.sorted_v[rb[i]] = vreg[i0 + i];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[slot]); // ring slot may be refilled
        VKRS_PHASE(7)
        prev_valid = valid;
    }
    if (timing && lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(dbg + 8 + k, tw[k]);
        atomicAdd(dbg + 4, (unsigned long long) tile_count);
    }
#undef VKRS_PHASE
    if (tile_count > 0) { // drain: the last tile of the segment
        named_bar_sync(bar_w, WORKERS);
        write_out((tile_count - 1) & 1, prev_valid);
    }
}

// =====================================================================================
// Helpers of the multi-GPU partition step.
// =====================================================================================
// totals[d] = sum over the segments of hist[g][d] (the 256 bucket counts of vkrs_partition).
__global__ void __launch_bounds__(RADIX)
segment_column_sum_kernel(const uint32_t *__restrict__ hist, uint32_t num_segments, uint32_t *__restrict__ totals) {
    uint32_t sum = 0;
    for (uint32_t g = 0; g < num_segments; ++g) sum += hist[(size_t) g * RADIX + threadIdx.x];
    totals[threadIdx.x] = sum;
}

// out[0] = min key, out[1] = max key (out pre-set to {0xFFFFFFFF, 0}); one coalesced read.
__global__ void __launch_bounds__(512)
key_range_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *out) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = ld_stream(keys + i);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

} // namespace vkrs
