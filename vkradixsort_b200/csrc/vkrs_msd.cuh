// vkrs_msd.cuh -- the keys-only "bucket" schedule: two most-significant-digit partition passes that
// do NOT have to be stable, then every 16-bit-prefix bucket is finished inside shared memory.
//
// Why: on B200 the stable digit pass (vkrs_segmented.cuh) is bound by the SM, not by HBM: ranking a
// key *stably* inside a warp costs 8 ballots + two nibble tables + two shuffles (~40 instructions
// per 32 keys) and three bank-conflicted shared-memory accesses.  Stability is what a
// least-significant-digit sort needs from every pass but its first.  A most-significant-digit split
// needs none: the order inside a bucket is irrelevant because the bucket is sorted completely
// afterwards.  Without stability a key's place in its tile is ONE shared-memory atomic
// (rank = atomicAdd(count[digit], 1)): 48 instead of 74 instructions, 17 instead of 23 shared-memory
// wavefronts per 32 keys (measured, profiles/README.md).  The result of a keys-only
// sort is unique, so the output is bit-identical to the reference's (testSort:
// multiradixsort/src/MultiRadixSort.cpp:148-161).  Key+payload sorts keep the stable LSD passes.
//
// Schedule for N keys (8-bit digits as in multi_radixsort.comp:12,100; the digits are taken from
// key - kmin, `top` = highest set bit of kmax - kmin, so only the OCCUPIED key range is sort work:
// leading zero bits -- the reference's 28-bit test keys, MultiRadixSort.cpp:126 --, a shared prefix,
// small signed integers around zero all cost nothing):
//   pass 1   digit ((key - kmin) >> s1) & 255, s1 = top-7 : piece histogram + unstable scatter  buf0 -> buf1
//   pass 2   digit ((key - kmin) >> s2) & 255, s2 = s1-8, inside each of the 256 buckets of pass 1
//                                                : piece histogram + unstable scatter  buf1 -> buf0
//   local    the (digit1, digit2) buckets (N / 65536 keys on average), batched into items of whole buckets
//            that fit a shared-memory buffer, are sorted in shared memory, in place in buf0: 4096
//            order-preserving bins ranked with atomics + a comparison fix-up inside the bins
//            (msd_local_tile_kernel below).
// A "piece" is the part of one bucket that lies inside one segment (a contiguous slab of the array
// owned by one worker group): the reference's "work group w owns nb*256 consecutive keys"
// (multi_radixsort_histograms.comp:43) cut at bucket boundaries, because a most-significant-digit
// pass must keep every key inside its bucket.  Row q of the histogram matrix belongs to piece q
// (multi_radixsort_histograms.comp:53-55), and the scatter's prologue turns the rows of ONE bucket
// into that piece's first output index per digit (multi_radixsort.comp:56-77, restricted to the bucket).
//
// The schedule depends on the key distribution: a (digit1, digit2) bucket larger than LOCAL_MAX keys
// cannot be finished in shared memory.  That is detected on the device while pass 2 runs (the bucket
// sizes fall out of its prologue) -- or already in pass 1, when a top-digit bucket exceeds
// 256 * LOCAL_MAX keys and pass 2 is then not run at all; the plan's `fallback` word is raised, the
// local sort does nothing and the four stable LSD passes that are enqueued behind it -- and otherwise
// exit at once -- sort buf0.  No host round trip, everything stays stream-ordered.
// Typed keys (XF != 0, KeyXform in vkrs_common.cuh): pass 1 reads the keys through the order-preserving
// map, the local sort (or the last fallback pass) writes them back through its inverse.
#pragma once
#include "vkrs_async.cuh"
#include "vkrs_common.cuh"

namespace vkrs {

// Device-resident control words of one bucket-schedule sort.
struct MsdPlan {
    uint32_t shift[2];      // digit shift of partition pass 1 / 2; shift[1] is also the number of low bits left to the local sort
    uint32_t fallback;      // != 0: some bucket is too large for the local sort -> the stable LSD passes run
    uint32_t recount;       // != 0: the pass-1 histogram was counted in the wrong digit window and is counted again
    uint32_t key_min;       // smallest key (gathered by the first histogram, with key_max)
    uint32_t max_sub;       // size of the largest (digit1, digit2) bucket pass 2 saw
    uint32_t num_pieces[2]; // pieces of pass 1 / 2 (NOT reset per sort: the pass-1 plan is cached per N)
    uint32_t key_max;       // largest key
    uint32_t skip_pass2;    // != 0: pass 1 already saw a bucket that is bound to overflow the local sort; pass 2 is not run
    uint32_t base;          // the digits are taken from key - base: the occupied key range starts at digit value 0
    uint32_t num_big;       // (digit1, digit2) buckets above LOCAL_MAX keys: finished by counting (msd_big_*), not in shared memory
    uint32_t big_items;     // work items (chunks of BIG_CHUNK keys) of their histograms
    uint32_t num_redo;      // items the bins path of the local sort could not take: sorted bucket by bucket by msd_local_redo_kernel
    uint32_t pad[2];
};

// Buckets too large for the shared-memory sort ("big" buckets: skewed keys -- half of the input under one 16-bit prefix,
// a few clusters, heavy duplicates).  All their keys agree above the low `shift[1]` <= 16 bits, so a big bucket is
// sorted by COUNTING: a histogram of the low bits (one global-memory counter per value, warp-aggregated atomics, any
// number of CTAs per bucket), an exclusive scan, and a fill that writes every value as often as it was counted -- no
// key moves.  4 B/key read + 4 B/key written, whatever the bucket's size.
constexpr uint32_t BIG_MAX = 1024;                 // big buckets per sort (more: the stable LSD passes take over)
constexpr uint32_t BIG_POOL_WORDS = 16u << 20;     // counters of all big buckets of one sort: 64 MB (256 buckets at 16 low bits)
constexpr uint32_t BIG_CHUNK = 65536;              // keys per histogram work item
constexpr uint32_t BIG_VCHUNK = 1024;              // values per fill work item
constexpr uint32_t BIG_VCS_STRIDE = 65536 / BIG_VCHUNK + 8; // first positions of a bucket's value chunks (+ end), per bucket
struct BigBucket {
    uint32_t lo, hi;        // the bucket's keys: positions [lo, hi) of the array
    uint32_t leaf;          // its (digit1, digit2) index: the keys are base + (leaf << low_bits) + [0, 2^low_bits)
    uint32_t first_item;    // its first histogram work item
};

constexpr int MSD_PLAN_THREADS = 1024;
constexpr int MSD_HIST_THREADS = 512;
constexpr int LOCAL_THREADS = 256;
constexpr int LOCAL_KPT = 16;
constexpr int LOCAL_MAX = LOCAL_THREADS * LOCAL_KPT; // largest bucket the local sort takes

// digit of a key: 8 bits of its distance from the base of the occupied key range
__device__ __forceinline__ uint32_t msd_digit(uint32_t key, uint32_t base, uint32_t shift) { return ((key - base) >> shift) & (RADIX - 1); }

// Exclusive scan of one value per thread over a 1024-thread block.  scratch = 33 uint32.
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *scratch, uint32_t *total_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t incl = warp_inclusive_scan(v, lane);
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = scratch[lane];
        const uint32_t wi = warp_inclusive_scan(w, lane);
        scratch[lane] = wi - w;
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    const uint32_t r = scratch[warp] + incl - v;
    if (total_out) *total_out = scratch[32];
    __syncthreads(); // scratch may be reused
    return r;
}

// Before the first histogram (only when the caller gave no key-span hint): a guess of the digit window from MSD_GUESS_SAMPLES
// keys -- 16 blocks of 1024 keys spread evenly over the array --, so that keys which do not fill the 32 bits -- the reference's own 28-bit test keys
// (MultiRadixSort.cpp:126), small non-negative integers -- are counted in their final window at once instead of being
// counted again (+60 us at 10^8 keys).  The guess is the window of the samples' range, pushed down by an eighth of a
// top-level bucket (the true smallest key lies a little below the smallest sample) or to 0; it is dropped if the
// largest sample would not fit it.  A wrong guess costs what no guess costs: msd_window_kernel checks the real range.
constexpr int MSD_GUESS_THREADS = 1024;
constexpr int MSD_GUESS_SAMPLES = 16384;
__host__ __device__ __forceinline__ bool msd_guess_window(uint32_t smin, uint32_t smax, uint32_t &shift0, uint32_t &base0) {
    const uint32_t span = smax - smin;
    uint32_t top = 0;
    while (top < 31 && (span >> (top + 1)) != 0) ++top;
    const uint32_t s1 = top >= 15u ? top - 7u : 8u;
    const uint32_t margin = 1u << (s1 - 3u);
    const uint32_t b0 = smin > margin ? smin - margin : 0u;
    if (((smax - b0) >> s1) >= (uint32_t) RADIX) return false;
    shift0 = s1;
    base0 = b0;
    return true;
}
// Start of a sort: flags down, and where the first histogram counts -- (base0, shift0) as the caller's key-span hint says
// (default: base 0, top byte), or, with `guess` set, the window guessed from the sample keys (see above; part of this
// kernel rather than one of its own: every kernel in front of the first histogram is a link of ~3 us in the launch chain).
template <int XF>
__global__ void __launch_bounds__(MSD_GUESS_THREADS)
msd_init_kernel(MsdPlan *plan, uint32_t shift0, uint32_t shift1, uint32_t base0, const uint32_t *__restrict__ keys, uint32_t n, uint32_t guess) {
    __shared__ uint32_t red[2][MSD_GUESS_THREADS / 32];
    const uint32_t tid = threadIdx.x;
    grid_dependency_wait();
    if (guess != 0 && n != 0) { // (uniform over the CTA)
        uint32_t lo = 0xFFFFFFFFu, hi = 0;
#pragma unroll
        for (int i = 0; i < MSD_GUESS_SAMPLES / MSD_GUESS_THREADS; ++i) {
            // 16 blocks of 1024 consecutive keys, the first at the start and the last at the end of the array: coalesced
            // reads from 16 places (16384 single keys spread over the array cost 9 us instead of 3: a page walk each)
            uint32_t idx = tid < n ? tid : n - 1u;
            if (n >= (uint32_t) MSD_GUESS_THREADS)
                idx += (uint32_t) (((uint64_t) i * (n - MSD_GUESS_THREADS)) / (MSD_GUESS_SAMPLES / MSD_GUESS_THREADS - 1));
            const uint32_t k = KeyXform<uint32_t, XF>::fwd(__ldg(keys + idx));
            lo = min(lo, k);
            hi = max(hi, k);
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((tid & 31) == 0) {
            red[0][tid >> 5] = lo;
            red[1][tid >> 5] = hi;
        }
        __syncthreads();
        if (tid < 32) {
            lo = __reduce_min_sync(0xffffffffu, red[0][tid]);
            hi = __reduce_max_sync(0xffffffffu, red[1][tid]);
            uint32_t gs = 0, gb = 0;
            if (msd_guess_window(lo, hi, gs, gb)) {
                shift0 = gs;
                shift1 = gs - 8u;
                base0 = gb;
            }
        }
    }
    if (tid == 0) {
        plan->shift[0] = shift0;
        plan->shift[1] = shift1;
        plan->base = base0;
        plan->fallback = 0;
        plan->recount = 0;
        plan->key_min = 0xFFFFFFFFu;
        plan->key_max = 0;
        plan->max_sub = 0;
        plan->skip_pass2 = 0;
        plan->num_big = 0;
        plan->big_items = 0;
        plan->num_redo = 0;
    }
}

// After the first histogram: the digits are taken from (key - smallest key), the two partition digits directly
// under the highest set bit of (largest - smallest key).  So only the OCCUPIED key range is sort work: leading zero
// bits (the reference's 28-bit test keys, MultiRadixSort.cpp:126), the shared prefix of one rank's key range after
// the multi-GPU exchange, small signed integers around zero (after the sign-flip map) all spread over the 65536
// buckets.  The histogram is only counted again when it was not already counted in that window.
__global__ void msd_window_kernel(MsdPlan *plan) {
    grid_dependency_wait();
    if (threadIdx.x == 0) {
        const uint32_t kmin = plan->key_min, kmax = plan->key_max;
        if (kmin <= kmax) {
            const uint32_t span = kmax - kmin;
            // all keys equal: nothing to sort; one bucket, no low bits left, no fallback
            const uint32_t top = span != 0 ? 31u - (uint32_t) __clz((int) span) : 0u;
            const uint32_t s1 = top >= 15u ? top - 7u : 8u;
            const uint32_t base0 = plan->base;
            const bool keep = base0 <= kmin && plan->shift[0] == s1 && ((kmax - base0) >> s1) < (uint32_t) RADIX;
            if (!keep) {
                plan->base = kmin;
                plan->shift[0] = s1;
                plan->shift[1] = s1 - 8u;
                plan->recount = 1;
            }
        }
    }
}

// =====================================================================================
// Pieces = segments cut at bucket boundaries.  One CTA; B <= 256 buckets, G <= 1024 segments.
//   bucket_start[b] (b = 1..B-1; entry 0 is taken as 0 and entry B as n; may be NULL when B == 1)
//   segment g = keys [g*seg_keys, (g+1)*seg_keys) of the array
// Output, all in position order: pieces[i] = (lo, hi, bucket, segment); seg_first[g] / bucket_first[b]
// = index of the first piece of segment g / bucket b (G+1 / B+1 entries).  sub_start (may be NULL):
// the B*256 + 1 starts of the (bucket, digit) runs the pass will produce -- this kernel fills the
// entries of EMPTY buckets and the end marker; the scatter fills the rest.
// =====================================================================================
__global__ void __launch_bounds__(MSD_PLAN_THREADS)
msd_plan_pieces_kernel(const uint32_t *__restrict__ bucket_start, uint32_t B, uint32_t n, uint32_t seg_keys, uint32_t G,
                       uint4 *__restrict__ pieces, uint32_t *__restrict__ seg_first, uint32_t *__restrict__ bucket_first,
                       uint32_t *__restrict__ num_pieces_out, uint32_t *__restrict__ sub_start,
                       const uint32_t *__restrict__ skip) {
    __shared__ uint32_t bs[RADIX + 1];
    __shared__ uint32_t ne[RADIX + 1]; // ne[b] = number of non-empty buckets among [0, b)
    __shared__ uint32_t scratch[33];
    const uint32_t tid = threadIdx.x;
    grid_dependency_wait();
    if (skip && *skip != 0) return;
    if (tid <= B) bs[tid] = tid == B ? n : (tid == 0 ? 0u : bucket_start[tid]);
    __syncthreads();
    const uint32_t nonempty = (tid < B && bs[tid + 1] > bs[tid]) ? 1u : 0u;
    const uint32_t ne_excl = block_exclusive_scan_1024(nonempty, scratch, nullptr);
    if (tid <= B) ne[tid] = ne_excl;
    __syncthreads();

    // first index i in [0, B] with bs[i] > x;  bs[0] = 0 <= x < n = bs[B]  =>  1 <= i <= B
    auto upper_bound = [&](uint32_t x) {
        uint32_t lo = 0, hi = B;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (bs[mid] > x) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    };
    uint32_t c = 0, b_lo = 0, b_hi = 0, seg_lo = 0, seg_hi = 0;
    if (tid < G) {
        const uint64_t lo64 = (uint64_t) tid * seg_keys, hi64 = lo64 + seg_keys;
        seg_lo = lo64 < n ? (uint32_t) lo64 : n;
        seg_hi = hi64 < n ? (uint32_t) hi64 : n;
        if (seg_lo < seg_hi) {
            b_lo = upper_bound(seg_lo) - 1;     // the bucket holding the segment's first key
            b_hi = upper_bound(seg_hi - 1) - 1; // ... and its last key
            c = ne[b_hi + 1] - ne[b_lo];
        }
    }
    uint32_t total = 0;
    const uint32_t first = block_exclusive_scan_1024(c, scratch, &total);
    if (tid < G) seg_first[tid] = first;
    if (tid == 0) {
        seg_first[G] = total;
        *num_pieces_out = total;
    }
    if (c > 0) {
        uint32_t i = first;
        for (uint32_t b = b_lo; b <= b_hi; ++b) {
            if (bs[b + 1] > bs[b]) {
                const uint32_t lo = bs[b] > seg_lo ? bs[b] : seg_lo, hi = bs[b + 1] < seg_hi ? bs[b + 1] : seg_hi;
                pieces[i++] = make_uint4(lo, hi, b, tid);
            }
        }
    }
    __syncthreads(); // the pieces (global memory) are visible to the whole CTA
    for (uint32_t i = tid; i < total; i += MSD_PLAN_THREADS) {
        const uint32_t pb = pieces[i].z;
        const int prev = i > 0 ? (int) pieces[i - 1].z : -1;
        for (int b = prev + 1; b <= (int) pb; ++b) bucket_first[b] = i;
        if (i == total - 1)
            for (uint32_t b = pb + 1; b <= B; ++b) bucket_first[b] = total;
    }
    if (total == 0 && tid <= B) bucket_first[tid] = 0;
    if (sub_start) {
        for (uint32_t idx = tid; idx < B * RADIX; idx += MSD_PLAN_THREADS) {
            const uint32_t b = idx >> RADIX_BITS;
            if (bs[b + 1] == bs[b]) sub_start[idx] = bs[b];
        }
        if (tid == 0) sub_start[B * RADIX] = n;
    }
}

// =====================================================================================
// hist[q][d] = #{keys of piece q whose digit is d}  (multi_radixsort_histograms.comp:31-56 with the
// work group's slab replaced by the piece).  One CTA per piece; 128-bit streaming loads; lane-private
// counter columns (bank == lane).  WITH_OR: also folds the smallest and the largest key into the plan.
// gate (may be NULL): the kernel only works if *gate != 0.
// =====================================================================================
template <bool WITH_OR, int XF = 0>
__global__ void __launch_bounds__(MSD_HIST_THREADS)
msd_piece_histogram_kernel(const uint32_t *__restrict__ keys, const uint4 *__restrict__ pieces,
                           const uint32_t *__restrict__ num_pieces, MsdPlan *plan, int pass, uint32_t *__restrict__ hist,
                           const uint32_t *__restrict__ gate) {
    __shared__ uint32_t cnt[RADIX * 32];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < RADIX * 32; i += MSD_HIST_THREADS) cnt[i] = 0;
    grid_dependency_wait();
    if (gate && *gate == 0) return;
    if (pass == 1 && plan->skip_pass2 != 0) return;
    if (blockIdx.x >= *num_pieces) return;
    const uint4 pc = pieces[blockIdx.x];
    const uint32_t shift = plan->shift[pass], kbase = plan->base;
    __syncthreads();
    uint32_t *my_col = cnt + lane;
    uint32_t acc_min = 0xFFFFFFFFu, acc_max = 0;
    auto count_key = [&](uint32_t raw) {
        const uint32_t k = KeyXform<uint32_t, XF>::fwd(raw); // typed keys: pass 1 reads them through the order-preserving map
        atomicAdd(my_col + msd_digit(k, kbase, shift) * 32, 1u);
        if (WITH_OR) {
            acc_min = min(acc_min, k);
            acc_max = max(acc_max, k);
        }
    };
    {
        const uint32_t *base = keys + pc.x;
        const uint32_t cnt_keys = pc.y - pc.x;
        uint32_t head = (uint32_t) (((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15) / sizeof(uint32_t));
        if (head > cnt_keys) head = cnt_keys;
        const uint32_t nvec = (cnt_keys - head) / 4;
        if ((uint32_t) tid < head) count_key(base[tid]);
        const uint4 *vbase = reinterpret_cast<const uint4 *>(base + head);
        uint32_t v = tid;
        for (; v + 3 * MSD_HIST_THREADS < nvec; v += 4 * MSD_HIST_THREADS) { // four 128-bit loads in flight
            uint4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = ld_stream(vbase + v + u * MSD_HIST_THREADS);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                count_key(a[u].x); count_key(a[u].y); count_key(a[u].z); count_key(a[u].w);
            }
        }
        for (; v < nvec; v += MSD_HIST_THREADS) {
            const uint4 a = ld_stream(vbase + v);
            count_key(a.x); count_key(a.y); count_key(a.z); count_key(a.w);
        }
        const uint32_t tail0 = head + nvec * 4;
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
    if (WITH_OR) {
        acc_min = __reduce_min_sync(0xffffffffu, acc_min);
        acc_max = __reduce_max_sync(0xffffffffu, acc_max);
        if (lane == 0 && acc_min <= acc_max) {
            atomicMin(&plan->key_min, acc_min);
            atomicMax(&plan->key_max, acc_max);
        }
    }
}

// =====================================================================================
// The unstable scatter.  Persistent CTA: GROUPS worker groups of WORKERS threads + one producer warp;
// group g of CTA c owns segment c*GROUPS + g and walks its pieces tile by tile:
//   producer   1-D TMA bulk copies (cp.async.bulk) of whole 16-byte aligned tiles into a two-slot
//              ring, completion on an mbarrier; a tile may reach past the piece (the keys outside
//              [lo, hi) are simply not looked at)
//   rank       r = atomicAdd(cnt[digit], 1): the key's place among the tile's keys of that digit
//              (this replaces the bin_flags bit matrix + popcounts of multi_radixsort.comp:97-118)
//   digit thr. tile-local exclusive scan of the 256 counts; the tile's global digit bases from running
//              per-digit offsets kept in registers (global_offsets[], multi_radixsort.comp:75-76,120-122);
//              at the first tile of a piece the offsets are rebuilt from the histogram rows of the
//              piece's bucket (the prologue, multi_radixsort.comp:56-77)
//   place      key -> sorted[excl[digit] + r] in shared memory
//   write-out  in tile order: a warp stores 128 B of consecutive positions; the write-out of tile j-1
//              overlaps the digit threads' work on tile j.
// UNIFORM_FAST: skew-aware ranking.  Per tile a warp looks at the first round of its chunk and picks one of three
// loops: plain atomics (the usual case; none of the tests below is compiled into it), "all one digit" (sorted or
// constant input: a uniform round is ranked with one atomic by lane 0 instead of 32 serialised ones), "hot digit"
// (a quarter or more of the round on one digit: those lanes share one atomic per round through a ballot).
// =====================================================================================
template <int WORKERS, int KPT>
struct MsdGroupSmem {
    static constexpr int TILE = WORKERS * KPT;
    alignas(128) uint32_t in[2][TILE];  // TMA destinations
    alignas(128) uint32_t sorted[TILE]; // tile in digit order, staged for the write-out
    uint32_t cnt[2][RADIX];             // digit counters of tile j (slot j & 1), then the digits' exclusive bases
    uint32_t bin_dst[2][RADIX];         // global start of the digit run minus its start in the tile
    uint32_t scan_scratch[8];
    alignas(8) uint64_t full[2], empty[2];
};

template <int WORKERS, int KPT, int GROUPS>
struct MsdSmem {
    using Group = MsdGroupSmem<WORKERS, KPT>;
    Group g[GROUPS];
};

template <int WORKERS, int KPT, int GROUPS, bool UNIFORM_FAST, int XF = 0>
__global__ void __launch_bounds__(GROUPS * WORKERS + 32, 1)
msd_scatter_kernel(const uint32_t *__restrict__ keys_in, uint32_t *__restrict__ keys_out, uint32_t n, MsdPlan *plan, int pass,
                   const uint4 *__restrict__ pieces, const uint32_t *__restrict__ seg_first,
                   const uint32_t *__restrict__ bucket_first, const uint32_t *__restrict__ bucket_start,
                   const uint32_t *__restrict__ hist, uint32_t *__restrict__ sub_start, uint32_t max_sub) {
    using Smem = MsdSmem<WORKERS, KPT, GROUPS>;
    using Group = typename Smem::Group;
    constexpr uint32_t TILE = Group::TILE;
    constexpr int WARPS = WORKERS / 32;
    constexpr int ALL_WORKERS = GROUPS * WORKERS;
    static_assert(WORKERS >= RADIX && WORKERS % 32 == 0, "one worker thread per digit is required");
    extern __shared__ __align__(128) unsigned char smem_raw_msd[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw_msd);

    const int tid = threadIdx.x, lane = tid & 31;
    const bool tma_ok = (reinterpret_cast<uintptr_t>(keys_in) & 15) == 0;

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
    if (tid < ALL_WORKERS) { // counters of the first tile (slot 0) start at zero
        const int grp0 = tid / WORKERS, g0 = tid - grp0 * WORKERS;
        if (g0 < RADIX) sm.g[grp0].cnt[0][g0] = 0;
    }
    grid_dependency_wait(); // plan, pieces, histogram rows and keys come from earlier kernels
    __syncthreads();
    // pass 2 is pointless once pass 1 has seen a bucket that must overflow the local sort; the flag is final
    // before this kernel starts (only pass 1 raises it), so all CTAs agree and buf0 keeps the original keys
    if (pass == 1 && plan->skip_pass2 != 0) return;

    // a tile goes through TMA when it is 16-byte aligned in global memory and lies inside the array
    auto tile_is_tma = [&](uint32_t tb) { return tma_ok && (uint64_t) tb + TILE <= (uint64_t) n; };

    if (tid >= ALL_WORKERS) {
        // ============================ producer warp (lane 0) ============================
        if (lane != 0) return;
        uint32_t q[GROUPS], q_end[GROUPS], t[GROUPS], nt[GROUPS], t0[GROUPS], jj[GROUPS];
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi) {
            const uint32_t seg = blockIdx.x * GROUPS + gi;
            q[gi] = seg_first[seg];
            q_end[gi] = seg_first[seg + 1];
            t[gi] = nt[gi] = t0[gi] = jj[gi] = 0;
        }
        bool any = true;
        while (any) {
            any = false;
#pragma unroll
            for (int gi = 0; gi < GROUPS; ++gi) {
                if (t[gi] == nt[gi]) { // next piece of this group's segment
                    if (q[gi] == q_end[gi]) continue;
                    const uint4 pc = pieces[q[gi]++];
                    t0[gi] = pc.x & ~3u;
                    nt[gi] = (pc.y - t0[gi] + TILE - 1) / TILE;
                    t[gi] = 0;
                }
                any = true;
                Group &s = sm.g[gi];
                const uint32_t slot = jj[gi] & 1, tb = t0[gi] + t[gi] * TILE;
                if (jj[gi] >= 2) mbar_wait_sleep(&s.empty[slot], ((jj[gi] - 2) >> 1) & 1, 64); // tile jj-2 fully consumed
                if (tile_is_tma(tb)) {
                    constexpr uint32_t kbytes = TILE * sizeof(uint32_t);
                    mbar_arrive_expect_tx(&s.full[slot], kbytes);
                    bulk_copy_g2s(s.in[slot], keys_in + tb, kbytes, &s.full[slot]);
                } else {
                    mbar_arrive(&s.full[slot]); // the workers copy this tile in themselves
                }
                ++t[gi];
                ++jj[gi];
            }
        }
        return;
    }

    // ==================================== workers ====================================
    const int grp = GROUPS == 1 ? 0 : tid / WORKERS;
    const int gtid = tid - grp * WORKERS, warp = gtid >> 5;
    Group &s = sm.g[grp];
    const uint32_t seg = blockIdx.x * GROUPS + grp;
    const uint32_t bar_w = 1 + grp, bar_d = 1 + GROUPS + grp; // this group's worker / digit named barriers
    const uint32_t shift = plan->shift[pass], kbase = plan->base;
    const uint32_t chunk0 = warp * (KPT * 32) + lane; // warp-striped: lane l reads chunk[i*32 + l]
    const bool is_digit_thread = gtid < RADIX;
    const uint32_t dgt = gtid, dwarp = dgt >> 5;
    const uint32_t q_begin = seg_first[seg], q_end = seg_first[seg + 1];

    uint32_t running_base = 0; // digit thread: where the next tile's run of its digit starts

    auto write_out = [&](uint32_t pslot, uint32_t count) {
        if (count == TILE) { // the usual case, without a bounds check (= a branch) per key
#pragma unroll
            for (int k = 0; k < KPT; ++k) {
                const uint32_t p = gtid + k * WORKERS;
                const uint32_t key = s.sorted[p];
                keys_out[s.bin_dst[pslot][msd_digit(key, kbase, shift)] + p] = key;
            }
        } else {
#pragma unroll
            for (int k = 0; k < KPT; ++k) {
                const uint32_t p = gtid + k * WORKERS;
                if (p < count) {
                    const uint32_t key = s.sorted[p];
                    keys_out[s.bin_dst[pslot][msd_digit(key, kbase, shift)] + p] = key;
                }
            }
        }
    };

    uint32_t jj = 0, prev_count = 0;
    for (uint32_t q = q_begin; q < q_end; ++q) {
        const uint4 pc = pieces[q]; // (lo, hi, bucket, segment)
        const uint32_t t0 = pc.x & ~3u;
        const uint32_t nt = (pc.y - t0 + TILE - 1) / TILE;
        for (uint32_t t = 0; t < nt; ++t, ++jj) {
            const uint32_t slot = jj & 1, par = (jj >> 1) & 1;
            const uint32_t tb = t0 + t * TILE;
            const uint32_t vlo = (pc.x > tb ? pc.x : tb) - tb;
            const uint32_t vhi = (pc.y < tb + TILE ? pc.y : tb + TILE) - tb;
            const uint32_t vcount = vhi - vlo; // keys of the piece inside this tile: positions [vlo, vhi)
            const bool full = vcount == TILE;
            const uint32_t *tin = s.in[slot];
            uint32_t *cnt = s.cnt[slot];
            mbar_wait(&s.full[slot], par);
            if (!tile_is_tma(tb)) {
                for (uint32_t p = gtid; p < TILE; p += WORKERS)
                    if (p - vlo < vcount) s.in[slot][p] = ld_stream(keys_in + tb + p);
                named_bar_sync(bar_w, WORKERS);
            }

            // ---- rank: one shared-memory atomic per key ----
            uint32_t rk[KPT];
            if (full) {
                // Skewed digits (sorted or constant input, half of the keys under one prefix): 32 atomics on one counter
                // are serialised.  The warp looks at the FIRST round of its chunk (once per tile): if at least a quarter
                // of its keys share one digit, the chunk is ranked by a loop that handles the lanes holding that "hot"
                // digit with ONE atomic per round (a ballot, the lowest such lane adds their number, the others take
                // their place from the ballot) and the rest with plain atomics.  That loop costs a vote and a shuffle
                // per key, which is why it is not the only one: compiled into the common loop the test cost 10 us per
                // pass.  A chunk that starts mixed and goes on skewed is ranked with plain atomics: slower, not wrong.
                uint32_t d_hot = RADIX; // no hot digit
                bool all_one = false;   // the whole first round shares one digit (sorted or constant input): a leaner loop
                if (UNIFORM_FAST) {
                    const uint32_t d0 = msd_digit(KeyXform<uint32_t, XF>::fwd(tin[chunk0]), kbase, shift);
                    const uint32_t same = __match_any_sync(0xffffffffu, d0);
                    const uint32_t most = __reduce_max_sync(0xffffffffu, (uint32_t) __popc(same));
                    all_one = most == 32u;
                    if (most >= 8u) {
                        const uint32_t holders = __ballot_sync(0xffffffffu, (uint32_t) __popc(same) == most);
                        d_hot = __shfl_sync(0xffffffffu, d0, __ffs((int) holders) - 1);
                    }
                }
                if (UNIFORM_FAST && all_one) {
#pragma unroll
                    for (int i = 0; i < KPT; ++i) {
                        const uint32_t d = msd_digit(KeyXform<uint32_t, XF>::fwd(tin[chunk0 + i * 32]), kbase, shift);
                        const uint32_t d_first = __shfl_sync(0xffffffffu, d, 0);
                        if (__all_sync(0xffffffffu, d == d_first)) { // one atomic by lane 0 instead of 32 serialised ones
                            uint32_t b = 0;
                            if (lane == 0) b = atomicAdd(&cnt[d], 32u);
                            rk[i] = __shfl_sync(0xffffffffu, b, 0) + lane;
                        } else {
                            rk[i] = atomicAdd(&cnt[d], 1u);
                        }
                    }
                } else if (UNIFORM_FAST && d_hot != (uint32_t) RADIX) {
                    const uint32_t lt_mask = lanemask_lt();
#pragma unroll
                    for (int i = 0; i < KPT; ++i) {
                        const uint32_t d = msd_digit(KeyXform<uint32_t, XF>::fwd(tin[chunk0 + i * 32]), kbase, shift);
                        const uint32_t hot = __ballot_sync(0xffffffffu, d == d_hot);
                        const int leader = __ffs((int) hot) - 1; // -1: no lane holds the hot digit in this round
                        uint32_t b = 0;
                        if ((int) lane == leader) b = atomicAdd(&cnt[d_hot], (uint32_t) __popc(hot));
                        b = __shfl_sync(0xffffffffu, b, leader < 0 ? 0 : leader);
                        rk[i] = d == d_hot ? b + (uint32_t) __popc(hot & lt_mask) : atomicAdd(&cnt[d], 1u);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < KPT; ++i)
                        rk[i] = atomicAdd(&cnt[msd_digit(KeyXform<uint32_t, XF>::fwd(tin[chunk0 + i * 32]), kbase, shift)], 1u);
                }
            } else {
#pragma unroll
                for (int i = 0; i < KPT; ++i) {
                    const uint32_t idx = chunk0 + i * 32;
                    rk[i] = 0;
                    if (idx - vlo < vcount) rk[i] = atomicAdd(&cnt[msd_digit(KeyXform<uint32_t, XF>::fwd(tin[idx]), kbase, shift)], 1u);
                }
            }
            named_bar_sync(bar_w, WORKERS); // (A) counts of tile jj final; sorted[] holds tile jj-1 completely

            // ---- digit threads: tile-local scan, this tile's global digit bases ----
            if (is_digit_thread) {
                const uint32_t total = cnt[dgt];
                const uint32_t incl = warp_inclusive_scan(total, lane);
                if (lane == 31) s.scan_scratch[dwarp] = incl;
                named_bar_sync(bar_d, RADIX);
                uint32_t warp_prefix = 0;
#pragma unroll
                for (int w = 0; w < RADIX / 32; ++w)
                    if (w < (int) dwarp) warp_prefix += s.scan_scratch[w];
                const uint32_t local_excl = warp_prefix + incl - total;
                cnt[dgt] = local_excl;        // place() adds the key's rank to this
                s.cnt[slot ^ 1][dgt] = 0;     // counters of tile jj+1 (its last readers passed barrier A)
                if (t == 0) {
                    // prologue of the piece (multi_radixsort.comp:56-77 over the rows of ONE bucket):
                    // first output index of this piece's keys with digit dgt
                    const uint32_t b = pc.z;
                    const uint32_t qf = bucket_first[b], ql = bucket_first[b + 1];
                    uint32_t below = 0, btotal = 0;
                    uint32_t q2 = qf;
                    for (; q2 + 8 <= ql; q2 += 8) { // eight independent loads in flight
                        uint32_t hh[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) hh[u] = __ldcg(hist + (size_t) (q2 + u) * RADIX + dgt);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            btotal += hh[u];
                            if (q2 + u < q) below += hh[u];
                        }
                    }
                    for (; q2 < ql; ++q2) {
                        const uint32_t hh = __ldcg(hist + (size_t) q2 * RADIX + dgt);
                        btotal += hh;
                        if (q2 < q) below += hh;
                    }
                    named_bar_sync(bar_d, RADIX); // everyone has read the tile scan's scratch
                    const uint32_t bincl = warp_inclusive_scan(btotal, lane);
                    if (lane == 31) s.scan_scratch[dwarp] = bincl;
                    named_bar_sync(bar_d, RADIX);
                    uint32_t bprefix = 0;
#pragma unroll
                    for (int w = 0; w < RADIX / 32; ++w)
                        if (w < (int) dwarp) bprefix += s.scan_scratch[w];
                    const uint32_t run_start = (bucket_start && b > 0 ? bucket_start[b] : 0u) + bprefix + bincl - btotal;
                    running_base = run_start + below;
                    if (sub_start && q == qf) { // the bucket's first piece publishes the run starts
                        sub_start[b * RADIX + dgt] = run_start;
                        if (max_sub) {
                            // the largest bucket the shared-memory sort will see (it sizes the items); a bucket above
                            // max_sub keys with low bits left is "big": finished by counting (msd_items_kernel lists them)
                            const uint32_t wmax = __reduce_max_sync(0xffffffffu, (btotal > max_sub && shift > 0) ? 0u : btotal);
                            if (lane == 0) atomicMax(&plan->max_sub, wmax);
                        }
                    }
                }
                s.bin_dst[slot][dgt] = running_base - local_excl;
                running_base += total;
            }
            // ---- write tile jj-1 out (overlaps the digit threads' work above) ----
            if (jj > 0) write_out(slot ^ 1, prev_count);
            named_bar_sync(bar_w, WORKERS); // (B) digit bases and bin_dst of tile jj ready; sorted[] free

            // ---- keys of tile jj to their place in the staging buffer ----
            if (full) {
#pragma unroll
                for (int i = 0; i < KPT; ++i) {
                    const uint32_t key = KeyXform<uint32_t, XF>::fwd(tin[chunk0 + i * 32]);
                    s.sorted[cnt[msd_digit(key, kbase, shift)] + rk[i]] = key;
                }
            } else {
#pragma unroll
                for (int i = 0; i < KPT; ++i) {
                    const uint32_t idx = chunk0 + i * 32;
                    if (idx - vlo < vcount) {
                        const uint32_t key = KeyXform<uint32_t, XF>::fwd(tin[idx]);
                        s.sorted[cnt[msd_digit(key, kbase, shift)] + rk[i]] = key;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.empty[slot]); // ring slot may be refilled
            prev_count = vcount;
        }
    }
    if (jj > 0) { // drain: the last tile of the segment
        named_bar_sync(bar_w, WORKERS);
        write_out((jj - 1) & 1, prev_count);
    }
}

// =====================================================================================
// Local sort.  After pass 2 the array is a sequence of 65536 (digit1, digit2) buckets in key order;
// what is left is to sort every bucket by its low bits.  Buckets are batched into ITEMS: item w = the
// buckets whose first key lies in window [w*W, (w+1)*W) of the array -- a contiguous range of whole
// buckets holding fewer than W + (largest bucket) keys.  W is picked on the device from the largest
// bucket pass 2 saw, so that an item fits the LT_CAP keys of a shared-memory buffer (lt_window()).  An
// item is sorted by the full key, which is the same thing because its buckets already are in order.
//
// msd_local_tile_kernel, one CTA per item at a time; the item's keys arrived in shared memory by one 1-D TMA bulk
// copy while the previous item was being sorted, and the sorted item goes back to the array by one bulk store while
// the next one is being counted (three key buffers rotate through the roles "this item's keys, then its
// sorted keys", "keys grouped by bin / the previous item on its way out", "copy target").  Two ways to sort an item:
//   bins     an order-preserving map of the item's key span onto LT_BINS bins, umulhi(key - base, mult) -- a
//            multiplier, not a shift, so that every bin is used whatever the span; one atomic per key to count,
//            a conflict-free 128-bit scan, one atomic to place; then position p ranks its key among the (few) keys
//            of its bin by comparison (local_tile_bins).  A bin of LT_BIN_LIMIT keys or more sends the item on.
//   buckets  the item's buckets one by one, two 8-bit passes in shared memory (local_bucket_sort): byte 0
//            unstable with atomics, byte 1 stable with the warp-private ballot ranking of the digit pass
//            (vkrs_tile.cuh / multi_radixsort.comp:97-122).  Always correct, several times slower.
// Nothing is written to global memory before a path has succeeded; the sorted item then goes back in
// place (store_item_bulk; store_item when the array is not 16-byte aligned).
// =====================================================================================
#ifndef VKRS_LT_CAP
#define VKRS_LT_CAP 7680
#endif
#ifndef VKRS_LT_CP_UNROLL
#define VKRS_LT_CP_UNROLL 16
#endif
#ifndef VKRS_LT_FIX_UNROLL
#define VKRS_LT_FIX_UNROLL 8
#endif
#ifndef VKRS_LT_BIN_BITS
#define VKRS_LT_BIN_BITS 12
#endif
#ifndef VKRS_LT_MIN_BLOCKS
#define VKRS_LT_MIN_BLOCKS 2
#endif
#ifndef VKRS_LT_THREADS
#define VKRS_LT_THREADS 512
#endif
constexpr int LT_THREADS = VKRS_LT_THREADS;
constexpr int LT_CAP = VKRS_LT_CAP; // keys per shared-memory buffer: the largest item the bins path takes
constexpr int LT_MIN_WINDOW = 256;
constexpr int LT_FIX_UNROLL = VKRS_LT_FIX_UNROLL; // positions per iteration of the fix-up loop
constexpr int LT_CP_UNROLL = VKRS_LT_CP_UNROLL;   // keys per iteration of the count and place loops
constexpr int LT_BIN_BITS = VKRS_LT_BIN_BITS;
#ifdef VKRS_LT_BINS
constexpr int LT_BINS = VKRS_LT_BINS; // (the bin map is a multiplication: any number of bins works)
#else
constexpr int LT_BINS = 1 << LT_BIN_BITS;
#endif
constexpr int LT_DESC = 64; // item descriptors loaded per batch (power of two, 4 * LT_DESC <= LT_THREADS)
constexpr int LT_BIN_LIMIT = 32; // power of two (the over-full test ORs the counts)
constexpr int LT_BPT = LT_BINS / LT_THREADS; // bins per thread in the scan: LT_BPT / 4 conflict-free 128-bit accesses
constexpr int LT_WORK_WORDS = LT_BINS > (LT_THREADS / 32) * RADIX ? LT_BINS : (LT_THREADS / 32) * RADIX;
static_assert(LT_BPT % 4 == 0 && LT_BPT >= 4, "whole 128-bit groups of bins per thread");
static_assert(LT_CAP >= LOCAL_MAX && LT_CAP < 65536, "part sums of the scan are packed two per word below; buffers double as the bucket path's a[] / b[]");

// Window of the item table for a largest bucket of max_sub keys: the largest multiple of 512 with
// window + max_sub <= LT_CAP - 4 (an item then fits a buffer), at least LT_MIN_WINDOW.
__host__ __device__ __forceinline__ uint32_t lt_window(uint32_t max_sub) {
    const uint32_t room = (uint32_t) LT_CAP - 4u; // the bins path takes items of up to LT_CAP - 4 keys
    if (max_sub + (uint32_t) LT_MIN_WINDOW >= room) return (uint32_t) LT_MIN_WINDOW;
    uint32_t w = (room - max_sub) & ~511u;
    if (w < (uint32_t) LT_MIN_WINDOW) w = (uint32_t) LT_MIN_WINDOW;
    return w > 16384u ? 16384u : w;
}

// TRAILING_SYNC = false: the caller guarantees a barrier before scratch is written again.
template <int THREADS, bool TRAILING_SYNC = true>
__device__ __forceinline__ uint32_t block_exclusive_scan_t(uint32_t v, uint32_t *scratch /* 33 */, uint32_t *total_out) {
    static_assert(THREADS % 32 == 0 && THREADS <= 1024, "whole warps");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t incl = warp_inclusive_scan(v, lane);
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < THREADS / 32 ? scratch[lane] : 0u;
        const uint32_t wi = warp_inclusive_scan(w, lane);
        scratch[lane] = wi - w;
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    const uint32_t r = scratch[warp] + incl - v;
    if (total_out) *total_out = scratch[32];
    if (TRAILING_SYNC) __syncthreads(); // scratch may be reused
    return r;
}

// item_first[w] = first bucket whose start is >= w*window, item_lo[w] = that bucket's start, for
// w = 0 .. ceil(n / window); the last entry is (num_sub, n).  One thread per bucket.
__global__ void __launch_bounds__(256)
msd_items_kernel(const uint32_t *__restrict__ sub_start, uint32_t num_sub, uint32_t n, uint32_t *__restrict__ item_first,
                 uint32_t *__restrict__ item_lo, uint32_t item_stride, MsdPlan *__restrict__ plan, BigBucket *__restrict__ big) {
    grid_dependency_wait();
    if (plan->fallback != 0) return;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num_sub) return;
    const uint32_t window = lt_window(plan->max_sub);
    const uint32_t num_items = (n + window - 1) / window;
    if (num_items + 1 > item_stride) return; // cannot happen: the table is sized for the smallest window
    const uint32_t s_j = __ldcg(sub_start + j);
    // ---- a big bucket: listed for the counting path, with one work item per BIG_CHUNK keys ----
    {
        const uint32_t e_j = __ldcg(sub_start + j + 1), low_bits = plan->shift[1];
        if (e_j - s_j > (uint32_t) LOCAL_MAX && low_bits > 0) {
            const uint32_t limit = min(BIG_MAX, BIG_POOL_WORDS >> low_bits);
            const uint32_t k = atomicAdd(&plan->num_big, 1u);
            if (k >= limit) {
                plan->fallback = 1; // too many of them: the stable LSD passes sort the array
            } else {
                const uint32_t chunks = (e_j - s_j + BIG_CHUNK - 1) / BIG_CHUNK;
                const uint32_t first = atomicAdd(&plan->big_items, chunks);
                big[k] = BigBucket{s_j, e_j, j, first};
            }
        }
    }
    // windows (w_prev, w_j] start inside the buckets before this one or at it: bucket j is the first whose start is
    // behind them.  After a bucket far larger than the window that is a long range: the warp fills it together.
    const int w_j = (int) (s_j / window);
    const int w_prev = j > 0 ? (int) (__ldcg(sub_start + j - 1) / window) : -1;
    const int last = (int) num_items - 1;
    auto fill = [&](int from, int to, uint32_t first, uint32_t lo) { // item_first / item_lo [from, to] <- (first, lo), warp-cooperative
        const int len = to - from + 1;
        if (len > 0 && len <= 4)
            for (int w = from; w <= to; ++w) {
                item_first[w] = first;
                item_lo[w] = lo;
            }
        uint32_t longer = __ballot_sync(0xffffffffu, len > 4);
        while (longer) {
            const int src = __ffs((int) longer) - 1;
            longer &= longer - 1;
            const int f0 = __shfl_sync(0xffffffffu, from, src), n0 = __shfl_sync(0xffffffffu, len, src);
            const uint32_t a = __shfl_sync(0xffffffffu, first, src), b = __shfl_sync(0xffffffffu, lo, src);
            for (int i = (int) (threadIdx.x & 31); i < n0; i += 32) {
                item_first[f0 + i] = a;
                item_lo[f0 + i] = b;
            }
        }
    };
    fill(w_prev + 1, w_j < last ? w_j : last, j, s_j);
    // windows in which no bucket starts (the tail of a large last bucket) ...
    const bool is_last = j == num_sub - 1;
    fill(is_last ? w_j + 1 : 1, is_last ? last : 0, num_sub, n);
    // ... and the end marker -- always: when the last buckets are empty and n is a multiple of the window, w_j is the
    // marker's own index and the range above is empty
    if (is_last) {
        item_first[last + 1] = num_sub;
        item_lo[last + 1] = n;
    }
}

struct LocalTileSmem {
    // three key buffers that rotate through the roles "this item's keys", "keys in sorted order" and
    // "destination of the copy of the next item's keys" (+ slack: 16-byte copy groups, neighbour reads)
    alignas(16) uint32_t buf[3][LT_CAP + 8];
    // work = work_m + 4.  bins path: bin counters, then running prefixes, work[-1] stays 0; bucket path: warp_cnt[16][256]
    alignas(16) uint32_t work_m[4 + LT_WORK_WORDS];
    uint32_t scratch[72];  // (40.. : per-warp minima / maxima of small_sort_kernel)
    uint32_t params[8]; // loop invariants that are only needed once per item: kept out of the registers
    uint32_t desc[LT_DESC][4]; // descriptors of the CTA's next items
    alignas(8) uint64_t copy_bar; // completion of the item copy in flight (one arrival per prefetch_item call)
    unsigned long long timers[16]; // phase timers of thread 0 (tuning aid, vkrs_debug_counters)
    unsigned long long t_last;
};

__device__ __forceinline__ void cp_async_4(uint32_t *smem_dst, const uint32_t *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t *smem_dst, const uint32_t *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Starts the copy of keys [lo, hi) into buf so that key lo + i lands at buf[(lo & 3) + i]: the whole 16-byte groups
// that lie inside the array with ONE 1-D TMA bulk copy issued by thread 0 (no LSU work at all: the cp.async version of
// this copy was 9 % of the kernel's shared-memory wavefronts and 6 % of its instructions), single keys (cp.async) for
// the rest -- the last keys of an array whose size is not a multiple of four, or an array that is not 16-byte aligned.
// n = size of the array.  Nothing is copied when the item does not fit a buffer (it goes to the redo list).  Every
// call completes one phase of `bar` (thread 0 arrives once, with the bulk copy's byte count if there is one); the
// consumer waits for cp_async_wait_all() AND that phase.  Called by all threads, after a barrier that follows the last
// read of buf.
__device__ __forceinline__ void prefetch_item(uint32_t *buf, const uint32_t *__restrict__ keys, uint32_t lo, uint32_t hi, uint32_t n,
                                              bool base_aligned, uint64_t *bar) {
    const uint32_t tid = threadIdx.x;
    uint32_t bulk_bytes = 0;
    const uint32_t a0 = lo & ~3u;
    if (hi > lo && hi - lo <= (uint32_t) LT_CAP) {
        if (base_aligned) {
            uint32_t a1 = (hi + 3u) & ~3u; // [a0, a1): 16-byte groups that lie inside the array
            if (a1 > (n & ~3u)) a1 = n & ~3u;
            if (a1 < a0) a1 = a0;
            bulk_bytes = (a1 - a0) * (uint32_t) sizeof(uint32_t);
            if (tid == 0) {
                bulk_store_wait_read(); // buf may have been the source of an earlier bulk store (store_item_bulk)
                for (uint32_t p = a1 > lo ? a1 : lo; p < hi; ++p) cp_async_4(&buf[p - a0], keys + p); // at most 3: the array's last keys
            }
        } else {
            for (uint32_t p = lo + tid; p < hi; p += LT_THREADS) cp_async_4(&buf[p - a0], keys + p);
        }
    }
    cp_async_commit();
    if (tid == 0) {
        if (bulk_bytes != 0) {
            mbar_arrive_expect_tx(bar, bulk_bytes);
            bulk_copy_g2s(buf, keys + a0, bulk_bytes, bar);
        } else {
            mbar_arrive(bar);
        }
    }
}

// Robust shared-memory sort of one bucket gk[0, cnt_keys), cnt_keys <= LOCAL_MAX, whose keys all lie in
// [bias, bias + 2^16): by the low 16 bits of key - bias (8 if !two_bytes).  All THREADS threads of the CTA call it.  a / b: LOCAL_MAX keys each; warp_cnt:
// [THREADS/32][256]; small_cnt: 256; scratch: 8.
template <int THREADS, int XF>
__device__ __forceinline__ void local_bucket_sort(uint32_t *__restrict__ gk, uint32_t cnt_keys, uint32_t bias, bool two_bytes, uint32_t *a,
                                                  uint32_t *b, uint32_t *warp_cnt, uint32_t *small_cnt, uint32_t *scratch) {
    constexpr int WARPS = THREADS / 32;
    constexpr int KPT = LOCAL_MAX / THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = lanemask_lt(), gt_mask = lanemask_gt();
    const DigitBitMasks bm(8);
    const LaneNibbleConsts lc(lane);
    uint32_t *my_cnt = warp_cnt + warp * RADIX;
    const uint32_t rounds = (cnt_keys + THREADS - 1) / THREADS;

    // ---- byte 0: load, count + rank with one atomic, scan, place into b[] ----
    uint32_t key[KPT], rk2[KPT / 2] = {}; // ranks < LOCAL_MAX: two per register
    auto set_rank = [&](int i, uint32_t r) {
        if (i & 1) rk2[i / 2] |= r << 16;
        else rk2[i / 2] = r;
    };
    auto get_rank = [&](int i) { return (i & 1) ? (rk2[i / 2] >> 16) : (rk2[i / 2] & 0xffffu); };
    if (tid < RADIX) small_cnt[tid] = 0;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        if (i < (int) rounds) {
            const uint32_t p = tid + i * THREADS;
            key[i] = p < cnt_keys ? ld_stream(gk + p) - bias : 0xFFFFFFFFu; // distance from the bucket's first key value: < 2^16
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        if (i < (int) rounds) {
            const uint32_t p = tid + i * THREADS;
            uint32_t r = 0;
            if (p < cnt_keys) r = atomicAdd(&small_cnt[key[i] & 255u], 1u);
            set_rank(i, r);
        }
    }
    __syncthreads();
    {
        const uint32_t total = tid < RADIX ? small_cnt[tid] : 0u;
        const uint32_t excl = block_exclusive_scan_256(total, scratch, nullptr);
        if (tid < RADIX) small_cnt[tid] = excl;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        if (i < (int) rounds) {
            const uint32_t p = tid + i * THREADS;
            if (p < cnt_keys) b[small_cnt[key[i] & 255u] + get_rank(i)] = key[i];
        }
    }
    __syncthreads();
    if (!two_bytes) {
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            if (i < (int) rounds) {
                const uint32_t p = tid + i * THREADS;
                if (p < cnt_keys) gk[p] = KeyXform<uint32_t, XF>::inv(b[p] + bias);
            }
        }
        __syncthreads();
        return;
    }

    // ---- byte 1, stable: warp w owns positions [w*rounds*32, (w+1)*rounds*32) of b[] ----
#pragma unroll
    for (int c = 0; c < RADIX / 32; ++c) my_cnt[lane + 32 * c] = 0;
    __syncwarp();
    const uint32_t p0 = warp * (rounds * 32) + lane;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        if (i < (int) rounds) {
            const uint32_t p = p0 + i * 32;
            // padding is all-ones: digit 255 and last in order, so it ranks behind every real key
            const uint32_t k = p < cnt_keys ? b[p] : 0xFFFFFFFFu;
            const uint32_t d = (k >> 8) & 255u;
            const uint32_t peers = match_key_table(k, d, bm, lc);
            const uint32_t r = my_cnt[d] + __popc(peers & lt_mask);
            if ((peers & gt_mask) == 0) my_cnt[d] = r + 1; // highest lane of the group
            key[i] = k;
            set_rank(i, r);
            __syncwarp();
        }
    }
    __syncthreads();
    {
        uint32_t total = 0;
        if (tid < RADIX) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) total += warp_cnt[w * RADIX + tid];
        }
        uint32_t running = block_exclusive_scan_256(total, scratch, nullptr);
        if (tid < RADIX) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = warp_cnt[w * RADIX + tid];
                warp_cnt[w * RADIX + tid] = running;
                running += c;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        if (i < (int) rounds) {
            const uint32_t p = p0 + i * 32;
            if (p < cnt_keys) a[my_cnt[(key[i] >> 8) & 255u] + get_rank(i)] = key[i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        if (i < (int) rounds) {
            const uint32_t p = tid + i * THREADS;
            if (p < cnt_keys) gk[p] = KeyXform<uint32_t, XF>::inv(a[p] + bias);
        }
    }
    __syncthreads(); // a[] / b[] / counters are free again
}

// Exclusive scan of TWO independent values per thread (v.x, v.y) over the block in one go.  scratch = 66 uint32.
// TRAILING_SYNC = false: the caller guarantees a barrier before scratch is written again.
template <int THREADS, bool TRAILING_SYNC = true>
__device__ __forceinline__ uint2 block_exclusive_scan2_t(uint2 v, uint32_t *scratch /* 66 */, uint2 *total_out) {
    static_assert(THREADS % 32 == 0 && THREADS <= 1024, "whole warps");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint2 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t tx = __shfl_up_sync(0xffffffffu, incl.x, o), ty = __shfl_up_sync(0xffffffffu, incl.y, o);
        if (lane >= o) {
            incl.x += tx;
            incl.y += ty;
        }
    }
    if (lane == 31) {
        scratch[warp] = incl.x;
        scratch[33 + warp] = incl.y;
    }
    __syncthreads();
    if (warp < 2) { // warp 0 scans the x sums, warp 1 the y sums
        uint32_t *sc = scratch + 33 * warp;
        const uint32_t w = lane < THREADS / 32 ? sc[lane] : 0u;
        const uint32_t wi = warp_inclusive_scan(w, lane);
        sc[lane] = wi - w;
        if (lane == 31) sc[32] = wi;
    }
    __syncthreads();
    const uint2 r = make_uint2(scratch[warp] + incl.x - v.x, scratch[33 + warp] + incl.y - v.y);
    if (total_out) *total_out = make_uint2(scratch[32], scratch[65]);
    if (TRAILING_SYNC) __syncthreads();
    return r;
}

// Bin map of an item that spans `span` key values: bin = umulhi(key - base, mult) with
// mult = floor(LT_BINS * 2^32 / span) -- order preserving, uses ALL bins whatever the span (a shift would leave up to
// half of them empty), always < LT_BINS.  lt_bin_mult() returns 0 when the span fits the bins ("exact mode": the caller
// then passes mult = 2^32 - 1 and base one less, which makes bin = key - base; equal bins are equal keys and nothing
// has to be fixed up).
__device__ __forceinline__ uint32_t lt_bin_mult(uint32_t num_buckets, uint32_t low_bits) {
    const uint64_t span = (uint64_t) num_buckets << low_bits;
    if (span <= (uint64_t) LT_BINS) return 0u;
    // floor(LT_BINS * 2^32 / span) to 21 bits, rounded DOWN (never above the exact quotient: the bin stays < LT_BINS;
    // a few of the topmost bins may go unused) -- a single-precision division instead of a 64-bit one per item
    const float q = __fdividef((float) LT_BINS * 4294967296.0f, (float) span) * (1.0f - 1.0f / 2097152.0f);
    return (uint32_t) q;
}

// Phase timers (tuning aid): thread 0 adds the clocks since the last mark to slot `i`.  Compiled in with -DVKRS_LT_TIMERS.
#ifdef VKRS_LT_TIMERS
#define LT_MARK(sm, i)                                        \
    do {                                                      \
        if (threadIdx.x == 0) {                               \
            const unsigned long long now__ = clock64();       \
            (sm).timers[i] += now__ - (sm).t_last;            \
            (sm).t_last = now__;                              \
        }                                                     \
    } while (0)
#else
#define LT_MARK(sm, i) do { } while (0)
#endif
// Bins path of one item: keys in `in[0, size)`, size <= LT_CAP - 4, umulhi(key - base, mult) < LT_BINS for every key;
// `gbuf` is a 16-byte aligned scratch buffer.  Leaves the key of final position p -- mapped back through the inverse of
// the typed-key transform XF -- in obuf[off + p] (obuf + off may be, and is, `in`) and returns false, or returns true
// (`in` untouched) when some bin is over-full.
//   count   one shared-memory atomic per key
//   scan    the bins' first positions
//   place   a second atomic on the bin's running position: the keys, grouped by bin, in gbuf[4 .. 4 + size)
//   fix-up  position p ranks its key among the keys of its bin by (key, offset inside the bin).  The first four
//           keys from the bin's start are compared without a branch: a key read past the bin's end belongs to a
//           later bin and is larger (the map is monotone), so it never counts.  Larger bins (rare: the expected
//           bin holds about one key) finish in a loop.
template <int XF = 0>
__device__ __forceinline__ bool local_tile_bins(LocalTileSmem &sm, const uint32_t *in, uint32_t *gbuf, uint32_t *obuf, uint32_t off,
                                                uint32_t size, uint32_t base, uint32_t mult, bool exact) {
    const int tid = threadIdx.x;
    uint32_t *cnt = sm.work_m + 4;
    uint4 *cv = reinterpret_cast<uint4 *>(cnt);
    uint32_t *grouped = gbuf + 4;
    constexpr int PARTS = LT_BPT / 4; // thread t owns bins [4t, 4t+4) of each of PARTS equal parts of the bin range
#pragma unroll
    for (int k = 0; k < PARTS; ++k) cv[k * LT_THREADS + tid] = make_uint4(0, 0, 0, 0);
    if (tid == 0) sm.work_m[3] = 0;
    __syncthreads();
    LT_MARK(sm, 2);

    // ---- count: one shared-memory reduction per key ----
#pragma unroll LT_CP_UNROLL
    for (uint32_t p = tid; p < size; p += LT_THREADS) atomicAdd(&cnt[__umulhi(in[p] - base, mult)], 1u);
    LT_MARK(sm, 3);
    __syncthreads();
    LT_MARK(sm, 4);

    // ---- scan: conflict-free 128-bit accesses; the part sums (< 2^16) are scanned two per word ----
    {
        static_assert(PARTS == 2 || PARTS == 4 || PARTS == 8, "two, four or eight parts");
        uint32_t sum[PARTS], orc = 0;
#pragma unroll
        for (int k = 0; k < PARTS; ++k) {
            const uint4 v = cv[k * LT_THREADS + tid];
            sum[k] = v.x + v.y + v.z + v.w;
            orc |= v.x | v.y | v.z | v.w;
        }
        uint32_t run[PARTS];
        if (PARTS == 2) {
            uint32_t total = 0;
            const uint32_t ex = block_exclusive_scan_t<LT_THREADS, false>(sum[0] | (sum[1] << 16), sm.scratch, &total); // a barrier follows below
            run[0] = ex & 0xffffu;
            run[1] = (ex >> 16) + (total & 0xffffu);
        } else {
            uint32_t part_base = 0;
#pragma unroll
            for (int k = 0; k < PARTS; k += 4) {
                uint2 total;
                const uint2 ex = block_exclusive_scan2_t<LT_THREADS, false>(
                    make_uint2(sum[k] | (sum[k + 1] << 16), sum[k + 2] | (sum[k + 3] << 16)), sm.scratch, &total);
                if (k + 4 < PARTS) __syncthreads(); // scratch is reused by the next round
                run[k] = part_base + (ex.x & 0xffffu);
                part_base += total.x & 0xffffu;
                run[k + 1] = part_base + (ex.x >> 16);
                part_base += total.x >> 16;
                run[k + 2] = part_base + (ex.y & 0xffffu);
                part_base += total.y & 0xffffu;
                run[k + 3] = part_base + (ex.y >> 16);
                part_base += total.y >> 16;
            }
        }
#pragma unroll
        for (int k = 0; k < PARTS; ++k) { // the counts are read again rather than kept in registers across the scan
            const uint4 v = cv[k * LT_THREADS + tid];
            uint4 o;
            o.x = run[k];
            o.y = o.x + v.x;
            o.z = o.y + v.y;
            o.w = o.z + v.z;
            cv[k * LT_THREADS + tid] = o;
        }
        // equal bins are equal keys in exact mode: nothing to fix up, any bin size is fine
        if (tid == 0) bulk_store_wait_read(); // gbuf may be the source of the previous item's bulk store: read before place writes it
        if (__syncthreads_or(!exact && orc >= (uint32_t) LT_BIN_LIMIT)) return true;
    }
    LT_MARK(sm, 5);

    // ---- place: a second atomic on the bin's running prefix hands out the positions; afterwards
    //      cnt[bin] = one past the last position of the bin, cnt[bin - 1] = its first.  (`gbuf` may still hold the
    //      previous item's output until the barrier behind the count: first touched here) ----
    if (tid < 4) grouped[size + tid] = 0xFFFFFFFFu; // the fix-up reads up to three keys past its bin: never smaller than a key
#pragma unroll LT_CP_UNROLL
    for (uint32_t p = tid; p < size; p += LT_THREADS) {
        const uint32_t k = in[p];
        grouped[atomicAdd(&cnt[__umulhi(k - base, mult)], 1u)] = k;
    }
    LT_MARK(sm, 6);
    __syncthreads();
    LT_MARK(sm, 7);
    if (exact) {
#pragma unroll 4
        for (uint32_t p = tid; p < size; p += LT_THREADS) obuf[off + p] = KeyXform<uint32_t, XF>::inv(grouped[p]);
        fence_proxy_async(); // the sorted item may leave through a bulk copy (store_item_bulk) after the next barrier
        return false;
    }

    // ---- fix-up, one position per thread: the bin's bounds from the counter array, the first four keys of the bin
    //      compared without a branch (a key read past the bin's end belongs to a later bin and is larger), a loop
    //      for larger bins ----
#pragma unroll LT_FIX_UNROLL
    for (uint32_t p = tid; p < size; p += LT_THREADS) {
        const uint32_t k = grouped[p];
        const uint32_t bin = __umulhi(k - base, mult);
        const uint32_t lo = cnt[(int) bin - 1], n = cnt[bin] - lo, d = p - lo;
        const uint32_t *g = grouped + lo;
        const uint64_t me = ((uint64_t) k << 32) | d;
        uint32_t r = lo;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) r += ((((uint64_t) g[j] << 32) | j) < me) ? 1u : 0u;
        if (n > 4) {
#pragma unroll 1
            for (uint32_t j = 4; j < n; ++j) r += ((((uint64_t) g[j] << 32) | j) < me) ? 1u : 0u;
        }
        obuf[off + r] = KeyXform<uint32_t, XF>::inv(k);
    }
    fence_proxy_async(); // the sorted item may leave through a bulk copy (store_item_bulk) after the next barrier
    LT_MARK(sm, 8);
    return false;
}

// Copies a sorted item (local_tile_bins) from shared memory back to its place: buf[(lo & 3) + p] -> keys[lo + p],
// 128-bit loads and stores where whole 16-byte groups of the array belong to the item.
template <int XF>
__device__ __forceinline__ void store_item(const uint32_t *buf, uint32_t *__restrict__ keys, uint32_t lo, uint32_t size, bool base_aligned) {
    const uint32_t tid = threadIdx.x;
    const uint32_t off = lo & 3u, a0 = lo - off, hi = lo + size;
    if (base_aligned) {
        const uint32_t groups = (off + size + 3u) >> 2;
        for (uint32_t g = tid; g < groups; g += LT_THREADS) {
            uint4 v = *reinterpret_cast<const uint4 *>(buf + 4u * g);
            v.x = KeyXform<uint32_t, XF>::inv(v.x);
            v.y = KeyXform<uint32_t, XF>::inv(v.y);
            v.z = KeyXform<uint32_t, XF>::inv(v.z);
            v.w = KeyXform<uint32_t, XF>::inv(v.w);
            const uint32_t at = a0 + 4u * g;
            if (at >= lo && at + 4u <= hi) {
                *reinterpret_cast<uint4 *>(keys + at) = v;
            } else { // the first / last group may be shared with the neighbouring items
                if (at + 0u >= lo && at + 0u < hi) keys[at + 0u] = v.x;
                if (at + 1u >= lo && at + 1u < hi) keys[at + 1u] = v.y;
                if (at + 2u >= lo && at + 2u < hi) keys[at + 2u] = v.z;
                if (at + 3u >= lo && at + 3u < hi) keys[at + 3u] = v.w;
            }
        }
    } else {
#pragma unroll 4
        for (uint32_t p = tid; p < size; p += LT_THREADS) keys[lo + p] = KeyXform<uint32_t, XF>::inv(buf[off + p]);
    }
}

// The same copy for untransformed keys in a 16-byte aligned array: the whole 16-byte groups of the item leave with ONE
// 1-D TMA bulk copy issued by thread 0 (no LDS / STG by the CTA), the up to three keys in front of and behind them with
// plain stores.  The writers of buf have executed fence_proxy_async() before the barrier in front of this call; thread 0
// waits for the copy's reads before buf is written again (local_tile_bins, prefetch_item) and for the copy itself at
// the end of the kernel.
__device__ __forceinline__ void store_item_bulk(const uint32_t *buf, uint32_t *__restrict__ keys, uint32_t lo, uint32_t size) {
    const uint32_t tid = threadIdx.x;
    const uint32_t a0 = lo & ~3u, hi = lo + size;
    const uint32_t g_lo = (lo + 3u) & ~3u, g_hi = hi & ~3u;
    if (g_hi > g_lo) {
        if (tid == 0) {
            bulk_copy_s2g(keys + g_lo, buf + (g_lo - a0), (g_hi - g_lo) * (uint32_t) sizeof(uint32_t));
            bulk_store_commit();
        }
        if (tid >= 32 && tid < 35) { // (a warp that does not issue the copy)
            const uint32_t p = lo + (tid - 32);
            if (p < g_lo) keys[p] = buf[p - a0];
        } else if (tid >= 36 && tid < 39) {
            const uint32_t p = g_hi + (tid - 36);
            if (p < hi) keys[p] = buf[p - a0];
        }
    } else {
        for (uint32_t p = lo + tid; p < hi; p += LT_THREADS) keys[p] = buf[p - a0];
    }
}

// Exclusive scan of a bucket's 2^low_bits counters, in place (one CTA; the counters live in L2), + vchunk_start[c] =
// first position of value c * BIG_VCHUNK, c = 0 .. ceil(V / BIG_VCHUNK) (the last entry = the bucket's size).
template <int THREADS>
__device__ __forceinline__ void big_scan_bucket(uint32_t *__restrict__ counters, uint32_t low_bits, uint32_t *__restrict__ vchunk_start,
                                                uint32_t *scratch /* 33 */) {
    constexpr uint32_t WARPS = THREADS / 32;
    const uint32_t V = 1u << low_bits, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t total = 0;
    if (V >= WARPS * 128u) {
        // One CTA does this while the others have nothing left to do: it must not crawl.  Warp w owns V / WARPS
        // consecutive counters and walks them 128 at a time (one 128-bit access per lane: coalesced, and the loads of a
        // pass do not depend on each other): a pass for the warp totals, a block scan of those, a pass that writes the
        // first positions.  (One thread per 64 consecutive counters, two dependent L2 round trips per counter, took
        // ~130 us for 65536 counters -- half of the histogram kernel's time for one bucket holding 5*10^7 keys.)
        const uint32_t per_warp = V / WARPS, steps = per_warp / 128u;
        uint4 *cv = reinterpret_cast<uint4 *>(counters + warp * per_warp) + lane;
        uint32_t sum = 0;
#pragma unroll 4
        for (uint32_t st = 0; st < steps; ++st) {
            const uint4 c = __ldcg(cv + 32u * st);
            sum += c.x + c.y + c.z + c.w;
        }
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) scratch[warp] = sum;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = lane < WARPS ? scratch[lane] : 0u;
            const uint32_t wi = warp_inclusive_scan(w, lane);
            scratch[lane] = wi - w;
            if (lane == 31) scratch[32] = wi;
        }
        __syncthreads();
        uint32_t run = scratch[warp];
        total = scratch[32];
#pragma unroll 4
        for (uint32_t st = 0; st < steps; ++st) {
            const uint4 c = __ldcg(cv + 32u * st);
            const uint32_t mine = c.x + c.y + c.z + c.w;
            const uint32_t incl = warp_inclusive_scan(mine, lane);
            const uint32_t excl = run + incl - mine;
            cv[32u * st] = make_uint4(excl, excl + c.x, excl + c.x + c.y, excl + c.x + c.y + c.z);
            const uint32_t v0 = warp * per_warp + st * 128u + 4u * lane;
            if ((v0 & (BIG_VCHUNK - 1)) == 0) vchunk_start[v0 / BIG_VCHUNK] = excl;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncthreads(); // scratch may be reused
    } else {
        const uint32_t per = (V + THREADS - 1) / THREADS;
        const uint32_t v0 = tid * per, v1 = min(V, v0 + per);
        uint32_t sum = 0;
        for (uint32_t v = v0; v < v1; ++v) sum += __ldcg(counters + v);
        uint32_t run = block_exclusive_scan_t<THREADS>(sum, scratch, &total);
        for (uint32_t v = v0; v < v1; ++v) {
            const uint32_t c = __ldcg(counters + v);
            counters[v] = run;
            if ((v & (BIG_VCHUNK - 1)) == 0) vchunk_start[v / BIG_VCHUNK] = run;
            run += c;
        }
    }
    if (tid == 0) vchunk_start[(V + BIG_VCHUNK - 1) / BIG_VCHUNK] = total;
}

// Histogram of the big buckets' low bits.  Persistent, one CTA per SM; CTA c takes a contiguous range of the work items
// (chunks of BIG_CHUNK keys; a bucket's chunks are consecutive items), counts into a shared-memory table of 65536
// 16-bit counters and flushes it to the bucket's global counters when the bucket changes: a value that occurs several
// times in the CTA's share costs one global atomic, not one per key (5 x fewer for one bucket holding half of 10^8
// keys).  A 16-bit counter that reaches 2^15 is spilled at once, so none can overflow.  The CTA that delivers a
// bucket's last chunk turns the bucket's counters into first positions (big_scan_bucket).
constexpr int BIG_HIST_THREADS = 1024;
template <int XF>
__global__ void __launch_bounds__(BIG_HIST_THREADS, 1)
msd_big_hist_kernel(const uint32_t *__restrict__ keys, const MsdPlan *__restrict__ plan, const BigBucket *__restrict__ big,
                    uint32_t *__restrict__ pool, uint32_t *__restrict__ big_done, uint32_t *__restrict__ big_vcs) {
    extern __shared__ __align__(16) uint32_t big_tab[]; // 32768 words = 65536 packed counters, + 40 words of scratch
    uint32_t *scratch = big_tab + 32768;
    grid_dependency_wait();
    if (plan->fallback != 0) return;
    const uint32_t big_items = plan->big_items, num_big = plan->num_big;
    if (big_items == 0) return;
    const uint32_t low_bits = plan->shift[1], mask = (1u << low_bits) - 1u, words = ((1u << low_bits) + 1) >> 1;
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (big_items + gridDim.x - 1) / gridDim.x;
    const uint32_t it0 = blockIdx.x * per, it1 = min(big_items, it0 + per);
    if (it0 >= it1) return;
    for (uint32_t i = tid; i < words; i += BIG_HIST_THREADS) big_tab[i] = 0;
    uint32_t cur = 0xFFFFFFFFu, cur_chunks = 0; // bucket being accumulated, chunks of it in the table
    BigBucket bk = {};
    uint32_t *counters = nullptr;
    __syncthreads();

    // the table goes to the bucket's global counters; the last deliverer of a bucket scans them
    auto flush = [&]() {
        for (uint32_t i = tid; i < words; i += BIG_HIST_THREADS) {
            const uint32_t w = big_tab[i];
            if (w != 0) {
                // one 64-bit reduction for the two 32-bit counters of values 2i and 2i+1 (the low one cannot carry: it
                // counts at most n < 2^30 keys): the flush is bound by the number of global atomics -- 148 CTAs x 65536
                // values for one bucket holding half of 10^8 keys, 160 of the kernel's 250 us with one atomic per value
                atomicAdd(reinterpret_cast<unsigned long long *>(counters + 2 * i), ((unsigned long long) (w >> 16) << 32) | (w & 0xffffu));
                big_tab[i] = 0;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const uint32_t chunks = (bk.hi - bk.lo + BIG_CHUNK - 1) / BIG_CHUNK;
            scratch[34] = atomicAdd(big_done + cur, cur_chunks) + cur_chunks == chunks ? 1u : 0u;
        }
        __syncthreads();
        if (scratch[34] != 0) {
            __threadfence();
            big_scan_bucket<BIG_HIST_THREADS>(counters, low_bits, big_vcs + (size_t) cur * BIG_VCS_STRIDE, scratch);
            if (tid == 0) big_done[cur] = 0; // for the next sort
        }
        __syncthreads();
    };

    for (uint32_t it = it0; it < it1; ++it) {
        // the bucket of this work item (the buckets were listed in any order: a linear search, uniform over the CTA)
        uint32_t k = cur, c = 0;
        if (cur != 0xFFFFFFFFu && it - bk.first_item < (bk.hi - bk.lo + BIG_CHUNK - 1) / BIG_CHUNK) {
            c = it - bk.first_item;
        } else {
            for (k = 0; k < num_big; ++k) {
                const BigBucket cand = big[k];
                if (it - cand.first_item < (cand.hi - cand.lo + BIG_CHUNK - 1) / BIG_CHUNK) {
                    c = it - cand.first_item;
                    break;
                }
            }
            if (cur != 0xFFFFFFFFu) flush();
            cur = k;
            cur_chunks = 0;
            bk = big[k];
            counters = pool + ((size_t) k << low_bits);
        }
        ++cur_chunks;
        const uint32_t value_base = plan->base + (bk.leaf << low_bits);
        const uint32_t lo = bk.lo + c * BIG_CHUNK, hi = min(bk.hi, lo + BIG_CHUNK);
        for (uint32_t p0 = lo + 4 * tid; p0 < hi; p0 += 4 * BIG_HIST_THREADS) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = p0 + u < hi ? ((__ldcs(keys + p0 + u) - value_base) & mask) : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (v[u] != 0xFFFFFFFFu) {
                    const uint32_t sh = (v[u] & 1u) << 4;
                    const uint32_t old = atomicAdd(&big_tab[v[u] >> 1], 1u << sh);
                    if (((old >> sh) & 0xffffu) == 0x7FFFu) { // this key made it 2^15: move that much to the global counter
                        atomicSub(&big_tab[v[u] >> 1], 0x8000u << sh);
                        // (64-bit like the flush: atomics of different sizes on one location do not mix)
                        atomicAdd(reinterpret_cast<unsigned long long *>(counters + (v[u] & ~1u)), 0x8000ull << (32u * (v[u] & 1u)));
                    }
                }
            }
        }
        __syncthreads();
    }
    flush();
}

// Fill: work item (bucket k, value chunk c).  A warp takes 32 consecutive values; a value counted once or twice is
// written by its lane, longer runs by the whole warp, 128 bytes per store.  The counters are left zero for the next sort.
template <int XF>
__global__ void __launch_bounds__(512)
msd_big_fill_kernel(uint32_t *__restrict__ keys, const MsdPlan *__restrict__ plan, const BigBucket *__restrict__ big,
                    uint32_t *__restrict__ pool, const uint32_t *__restrict__ vchunk_start_all) {
    grid_dependency_wait();
    if (plan->fallback != 0) return;
    const uint32_t num_big = plan->num_big;
    if (num_big == 0) return;
    const uint32_t low_bits = plan->shift[1], V = 1u << low_bits;
    const uint32_t vchunks = (V + BIG_VCHUNK - 1) / BIG_VCHUNK, items = num_big * vchunks;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    for (uint32_t it = blockIdx.x; it < items; it += gridDim.x) {
        const uint32_t k = it / vchunks, c = it - k * vchunks;
        const BigBucket bk = big[k];
        uint32_t *counters = pool + ((size_t) k << low_bits);
        const uint32_t *vcs = vchunk_start_all + (size_t) k * BIG_VCS_STRIDE;
        const uint32_t value_base = plan->base + (bk.leaf << low_bits);
        const uint32_t vlo = c * BIG_VCHUNK, vhi = min(V, vlo + BIG_VCHUNK);
        const uint32_t chunk_end = __ldcg(vcs + c + 1); // first position of the value behind this chunk
        for (uint32_t vb = vlo + 32 * warp; vb < vhi; vb += 32 * warps) {
            const uint32_t v = vb + lane;
            const bool valid = v < vhi;
            const uint32_t start = valid ? __ldcg(counters + v) : chunk_end;
            uint32_t next = __shfl_down_sync(0xffffffffu, start, 1); // first position of value v + 1
            if (valid) {
                if (v + 1 >= vhi) next = chunk_end;
                else if (lane == 31) next = __ldcg(counters + v + 1);
            }
            const uint32_t cnt = valid ? next - start : 0u;
            const uint32_t key = KeyXform<uint32_t, XF>::inv(value_base + v);
            if (cnt >= 1 && cnt <= 2) {
                keys[bk.lo + start] = key;
                if (cnt == 2) keys[bk.lo + start + 1] = key;
            }
            uint32_t longer = __ballot_sync(0xffffffffu, cnt > 2);
            while (longer) {
                const int src = __ffs((int) longer) - 1;
                longer &= longer - 1;
                const uint32_t s0 = __shfl_sync(0xffffffffu, start, src), n0 = __shfl_sync(0xffffffffu, cnt, src);
                const uint32_t kk = __shfl_sync(0xffffffffu, key, src);
                for (uint32_t i = lane; i < n0; i += 32) keys[bk.lo + s0 + i] = kk;
            }
        }
        __syncthreads(); // every start of the chunk has been read (a warp's last lane reads its neighbour's first)
        for (uint32_t v = vlo + threadIdx.x; v < vhi; v += blockDim.x) counters[v] = 0;
    }
}

// XF != 0 (typed keys): the keys in the array are the transformed ones; every key is written back through the inverse map.
template <int XF>
__global__ void __launch_bounds__(LT_THREADS, VKRS_LT_MIN_BLOCKS)
msd_local_tile_kernel(uint32_t *__restrict__ keys, const uint32_t *__restrict__ sub_start,
                      const uint32_t *__restrict__ item_first, const uint32_t *__restrict__ item_lo, uint32_t n,
                      MsdPlan *__restrict__ plan, uint32_t use_bins /* 0: per-bucket path only (tests) */,
                      unsigned long long *__restrict__ timers_out, uint32_t *__restrict__ redo) {
    extern __shared__ __align__(128) unsigned char smem_raw_tile[];
    LocalTileSmem &sm = *reinterpret_cast<LocalTileSmem *>(smem_raw_tile);
    const int tid = threadIdx.x;
    grid_dependency_wait();
    if (plan->fallback != 0) return;
    const uint32_t low_bits = plan->shift[1];
    if (low_bits == 0) { // nothing left to sort; typed keys still have to be mapped back
        if (XF != 0)
            for (uint32_t i = blockIdx.x * LT_THREADS + tid; i < n; i += gridDim.x * LT_THREADS) keys[i] = KeyXform<uint32_t, XF>::inv(keys[i]);
        return;
    }
    if (tid == 0) {
        sm.params[0] = plan->base; // bucket j holds the keys base + (j << low_bits) + [0, 2^low_bits)
        sm.params[1] = use_bins;
        mbar_init(&sm.copy_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t window = lt_window(plan->max_sub);
    const uint32_t num_items = (n + window - 1) / window;

    uint32_t w = blockIdx.x;
    if (w >= num_items) return;
    // Keys one item ahead of the sort; the descriptors (lo, hi, first bucket, end bucket) of the CTA's next 64 items are
    // loaded together every 64 items, right in front of a barrier, so that no item waits for a global load.
    // (Measured alternatives, both slower than even a load per item: a shared-memory ring filled by a plain load left in
    // flight across the item -- every later instruction that shares its scoreboard slot waits for it, and the warp that
    // issued it arrives late at each barrier, +75 us -- or by cp.async together with the keys, +50 us.)
    uint32_t lo = __ldcg(item_lo + w), hi = __ldcg(item_lo + w + 1);
    uint32_t j0 = __ldcg(item_first + w), j1 = __ldcg(item_first + w + 1);
    uint32_t b_in = 0, b_sorted = 1, b_next = 2; // roles of the three key buffers
    const bool base_aligned = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
    prefetch_item(sm.buf[b_in], keys, lo, hi, n, base_aligned, &sm.copy_bar);
    if (tid == 0) sm.params[2] = lt_bin_mult(j1 > j0 ? j1 - j0 : 1u, low_bits); // the bin map of an item is computed one item ahead
#ifdef VKRS_LT_TIMERS
    if (tid == 0) {
        for (int i = 0; i < 16; ++i) sm.timers[i] = 0;
        sm.t_last = clock64();
    }
#endif
    uint32_t mslot = 0, iter = 0;
    uint32_t pend_lo = 0, pend_size = 0; // the previous item, sorted, waits in buf[b_sorted] for its copy back to the array
    for (; w < num_items; w += gridDim.x, ++iter) {
        if ((iter & (LT_DESC - 1)) == 0 && tid < 4 * LT_DESC) { // the descriptors of the items behind this one
            const uint32_t item = w + ((tid >> 2) + 1) * gridDim.x;
            sm.desc[tid >> 2][tid & 3] = item < num_items ? __ldcg(((tid & 3) < 2 ? item_lo : item_first) + item + (tid & 1)) : 0u;
        }
        const uint32_t size = hi - lo;
        cp_async_wait_all();
        mbar_wait(&sm.copy_bar, iter & 1u);
        LT_MARK(sm, 9);
        __syncthreads(); // this item's keys are in buf[b_in]; buf[b_next] and the work area are free; buf[b_sorted] = previous item, sorted
        LT_MARK(sm, 0);
        const uint32_t *nd = sm.desc[iter & (LT_DESC - 1)];
        const uint32_t nlo = nd[0], nhi = nd[1], nj0 = nd[2], nj1 = nd[3];
        // ---- start the copy of the next item: it has the whole of this item's sort to land ----
        prefetch_item(sm.buf[b_next], keys, nlo, nhi, n, base_aligned, &sm.copy_bar);
        if (tid == 0) sm.params[2 + (mslot ^ 1)] = lt_bin_mult(nj1 > nj0 ? nj1 - nj0 : 1u, low_bits);
        // ---- the previous item goes back to the array while this one is counted (buf[b_sorted] is first written two barriers on) ----
        if (pend_size != 0) {
            if (base_aligned) store_item_bulk(sm.buf[b_sorted], keys, pend_lo, pend_size);
            else store_item<0>(sm.buf[b_sorted], keys, pend_lo, pend_size, base_aligned);
        }
        pend_size = 0;
        LT_MARK(sm, 1);
        if (size > 1) {
            bool todo = true;
            if (size <= (uint32_t) LT_CAP) {
                // the item's buckets are j0 .. j1-1 and a key of bucket j is kbase + (j << low_bits) + its low bits
                const uint32_t base = sm.params[0] + (j0 << low_bits);
                uint32_t *in = sm.buf[b_in] + (lo & 3u);
                if (sm.params[1] != 0 && size <= (uint32_t) LT_CAP - 4u) {
                    const uint32_t mult = sm.params[2 + mslot];
                    const bool exact = mult == 0;
                    todo = local_tile_bins<XF>(sm, in, sm.buf[b_sorted], sm.buf[b_in], lo & 3u, size, exact ? base - 1u : base, exact ? 0xFFFFFFFFu : mult, exact);
                }
                if (!todo) {
                    pend_lo = lo;
                    pend_size = size;
                }
            }
            if (todo && tid == 0) redo[atomicAdd(&plan->num_redo, 1u)] = w; // rare: msd_local_redo_kernel sorts it bucket by bucket
        } else if (XF != 0 && size == 1 && tid == 0) {
            keys[lo] = KeyXform<uint32_t, XF>::inv(keys[lo]);
        }
        LT_MARK(sm, 11);
        lo = nlo;
        hi = nhi;
        j0 = nj0;
        j1 = nj1;
        mslot ^= 1;
        const uint32_t t = b_in; // rotate: next keys <- prefetched, sorted scratch <- old keys (= the sorted item), prefetch target <- old scratch
        b_in = b_next;
        b_next = b_sorted;
        b_sorted = t;
    }
    if (pend_size != 0) {
        __syncthreads();
        if (base_aligned) store_item_bulk(sm.buf[b_sorted], keys, pend_lo, pend_size);
        else store_item<0>(sm.buf[b_sorted], keys, pend_lo, pend_size, base_aligned);
    }
    if (tid == 0) bulk_store_wait_all();
#ifdef VKRS_LT_TIMERS
    if (tid == 0 && timers_out) {
        sm.timers[10] += 1; // CTAs
        for (int i = 0; i < 16; ++i) atomicAdd(&timers_out[i], sm.timers[i]);
    }
#endif
    cp_async_wait_all();
}

// =====================================================================================
// Small sorts in ONE launch and a handful of barriers: up to LT_CAP - 4 keys are one "item" of the local sort -- load,
// smallest / largest key, the multiplicative bin map over that span, count, scan, place, fix-up, store.  (The
// reference's single_radixsort.comp runs four digit passes with three barriers per 256 keys inside one work group;
// the restated kernel, single_sort_kernel, takes 35 us for 1000 keys -- launch latency and ~50 barriers.)  An over-full
// bin (a few distinct values far apart) sends the keys through a bitonic network in shared memory instead.
// =====================================================================================
static_assert(2 * (LT_CAP + 8) >= 8192 && LT_CAP <= 8192, "the bitonic fallback of small_sort_kernel pads to at most 8192 keys in two buffers");
__global__ void __launch_bounds__(LT_THREADS, 1)
small_sort_kernel(uint32_t *__restrict__ keys, uint32_t n) {
    extern __shared__ __align__(128) unsigned char smem_raw_small[];
    LocalTileSmem &sm = *reinterpret_cast<LocalTileSmem *>(smem_raw_small);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    grid_dependency_wait();
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0;
    for (uint32_t p = tid; p < n; p += LT_THREADS) {
        const uint32_t k = keys[p];
        sm.buf[0][p] = k;
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) {
        sm.scratch[40 + warp] = kmin;
        sm.scratch[56 + warp] = kmax;
    }
    __syncthreads();
    for (int w = 0; w < LT_THREADS / 32; ++w) {
        kmin = min(kmin, sm.scratch[40 + w]);
        kmax = max(kmax, sm.scratch[56 + w]);
    }
    const uint64_t span = (uint64_t) (kmax - kmin) + 1u;
    const bool exact = span <= (uint64_t) LT_BINS;
    uint32_t mult = 0xFFFFFFFFu;
    if (!exact) mult = (uint32_t) (__fdividef((float) LT_BINS * 4294967296.0f, (float) span) * (1.0f - 1.0f / 2097152.0f));
    const bool todo = local_tile_bins(sm, sm.buf[0], sm.buf[1], sm.buf[0], 0u, n, exact ? kmin - 1u : kmin, mult, exact);
    if (todo) {
        // an over-full bin (a few distinct values far apart): a bitonic network over the keys, padded with all-ones to
        // a power of two, in the first two buffers taken as one (2 * (LT_CAP + 8) >= 8192 words).  Rare and small.
        uint32_t *a = sm.buf[0];
        uint32_t m = 1;
        while (m < n) m <<= 1;
        __syncthreads();
        for (uint32_t p = n + tid; p < m; p += LT_THREADS) a[p] = 0xFFFFFFFFu;
        __syncthreads();
        for (uint32_t k = 2; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = tid; i < m; i += LT_THREADS) {
                    const uint32_t l = i ^ j;
                    if (l > i) {
                        const uint32_t x = a[i], y = a[l];
                        if (((i & k) == 0) == (x > y)) {
                            a[i] = y;
                            a[l] = x;
                        }
                    }
                }
                __syncthreads();
            }
    }
    __syncthreads();
    store_item<0>(sm.buf[0], keys, 0u, n, (reinterpret_cast<uintptr_t>(keys) & 15) == 0);
}

// =====================================================================================
// The items the bins path could not take (an over-full bin: heavy duplicates; an item that does not fit a buffer
// because a big bucket lies in its window; VKRS_LOCAL_BINS=0), bucket by bucket with local_bucket_sort.  A kernel of
// its own: inlined into msd_local_tile_kernel this rarely used path cost the hot loop its registers (64 with 76 bytes
// of spills against 52 without: 555 -> 463 us for 10^8 keys).  Exits at once when the list is empty.
// =====================================================================================
struct LocalRedoSmem {
    alignas(16) uint32_t a[LOCAL_MAX + 8], b[LOCAL_MAX + 8];
    uint32_t warp_cnt[(LT_THREADS / 32) * RADIX];
    uint32_t small_cnt[RADIX];
    uint32_t cand_lo[LT_THREADS], cand_hi[LT_THREADS];
    uint32_t scratch[40];
};

template <int XF>
__global__ void __launch_bounds__(LT_THREADS)
msd_local_redo_kernel(uint32_t *__restrict__ keys, const uint32_t *__restrict__ sub_start, const uint32_t *__restrict__ item_first,
                      const MsdPlan *__restrict__ plan, const uint32_t *__restrict__ redo) {
    extern __shared__ __align__(128) unsigned char smem_raw_redo[];
    LocalRedoSmem &sm = *reinterpret_cast<LocalRedoSmem *>(smem_raw_redo);
    const uint32_t tid = threadIdx.x;
    grid_dependency_wait();
    if (plan->fallback != 0) return;
    const uint32_t num_redo = plan->num_redo, low_bits = plan->shift[1], kbase = plan->base;
    for (uint32_t r = blockIdx.x; r < num_redo; r += gridDim.x) {
        const uint32_t w = redo[r];
        const uint32_t j0 = __ldcg(item_first + w), j1 = __ldcg(item_first + w + 1);
        // the item's buckets one by one (empty and one-key buckets are skipped in chunks; buckets above LOCAL_MAX
        // keys are not touched: they are "big" and sorted by counting)
        for (uint32_t jb = j0; jb < j1; jb += LT_THREADS) {
            const uint32_t j = jb + tid;
            uint32_t blo = 0, bhi = 0;
            if (j < j1) {
                blo = __ldcg(sub_start + j);
                bhi = __ldcg(sub_start + j + 1);
            }
            sm.cand_lo[tid] = blo;
            sm.cand_hi[tid] = bhi;
            if (__syncthreads_or(bhi - blo > (XF != 0 ? 0u : 1u)) == 0) continue;
            const uint32_t chunk = j1 - jb < (uint32_t) LT_THREADS ? j1 - jb : (uint32_t) LT_THREADS;
            for (uint32_t c = 0; c < chunk; ++c) {
                const uint32_t clo = sm.cand_lo[c], chi = sm.cand_hi[c];
                if (XF != 0 && chi - clo == 1 && tid == 0) keys[clo] = KeyXform<uint32_t, XF>::inv(keys[clo]);
                if (chi - clo > 1 && chi - clo <= (uint32_t) LOCAL_MAX)
                    local_bucket_sort<LT_THREADS, XF>(keys + clo, chi - clo, kbase + ((jb + c) << low_bits), low_bits > 8, sm.a, sm.b, sm.warp_cnt,
                                                      sm.small_cnt, sm.scratch);
            }
            __syncthreads(); // cand_lo / cand_hi are rewritten by the next chunk
        }
    }
}

} // namespace vkrs
