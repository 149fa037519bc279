// vkrs_kernels.cuh -- all __global__ kernels of the B200 radix sort.
//
// Default whole sort:                 vkrs_segmented.cuh (segment_histogram_kernel + segmented_scatter_kernel)
// Single-sweep variants ("simple"):   global_histogram_kernel -> NUM_PASSES x onesweep_pass_kernel
//                        ("pipe"):    global_histogram_kernel -> NUM_PASSES x onesweep_pipelined_kernel (vkrs_pipeline.cuh)
// Staged path (per-stage C-ABI):      staged_histograms_kernel | staged_colsum/chunkscan/offsets + staged_scatter_kernel
// Single-workgroup path:              single_sort_kernel
#pragma once
#include "vkrs_tile.cuh"

namespace vkrs {

// =====================================================================================
// Fused path, kernel 1: the NUM_PASSES x 256 digit histograms of the whole input in one
// read of the keys (4 B/key), exclusive-scanned in place by the last CTA to finish.
//
// Shared-memory layout is bank-conflict-free by construction: every lane owns a private
// column.  Counter (pass p, digit d, lane l) is the 16-bit half (d & 1) of word
//   hist[p][d >> 1][l]            -> bank == lane for every access, no two lanes of a warp
// ever touch the same bank, so each atomic instruction is a single shared-memory wavefront
// whatever the key distribution (all-equal keys included).  A CTA may feed at most
// 65535 keys into one lane column before the halves could carry; the host sizes the grid
// so that keys_per_cta / 32 < 65536.
// =====================================================================================
constexpr int HIST_THREADS = 512;
constexpr uint32_t HIST_MAX_KEYS_PER_CTA = 1u << 20;

template <typename KeyT, int NUM_PASSES>
__global__ void __launch_bounds__(HIST_THREADS)
global_histogram_kernel(const KeyT *__restrict__ keys, uint32_t n, uint32_t keys_per_cta, uint32_t *ghist,
                        uint32_t *done_counter) {
    extern __shared__ uint32_t hist_smem[]; // [NUM_PASSES][128][32]
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < NUM_PASSES * 128 * 32; i += HIST_THREADS) hist_smem[i] = 0;
    __syncthreads();

    const uint64_t lo = (uint64_t) blockIdx.x * keys_per_cta;
    const uint64_t hi = (lo + keys_per_cta < n) ? lo + keys_per_cta : n;

    auto count_key = [&](KeyT k) {
#pragma unroll
        for (int p = 0; p < NUM_PASSES; ++p) {
            const uint32_t d = static_cast<uint32_t>(k >> (8 * p)) & 255u;
            atomicAdd(&hist_smem[(p * 128 + (d >> 1)) * 32 + lane], (d & 1u) ? 0x10000u : 1u);
        }
    };

    if (lo < hi) {
        constexpr int VEC = 16 / sizeof(KeyT); // keys per 128-bit load
        const KeyT *base = keys + lo;
        const uint64_t count = hi - lo;
        // scalar head up to 16-byte alignment, 128-bit body, scalar tail
        uint64_t head = ((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15) / sizeof(KeyT);
        if (head > count) head = count;
        const uint64_t nvec = (count - head) / VEC;
        if (tid < head) count_key(base[tid]);
        const uint4 *vbase = reinterpret_cast<const uint4 *>(base + head);
        uint64_t v = tid;
        // two independent 128-bit loads in flight per thread
        for (; v + HIST_THREADS < nvec; v += 2 * HIST_THREADS) {
            const uint4 a = ld_stream(vbase + v);
            const uint4 b = ld_stream(vbase + v + HIST_THREADS);
            if (sizeof(KeyT) == 4) {
                count_key((KeyT) a.x); count_key((KeyT) a.y); count_key((KeyT) a.z); count_key((KeyT) a.w);
                count_key((KeyT) b.x); count_key((KeyT) b.y); count_key((KeyT) b.z); count_key((KeyT) b.w);
            } else {
                count_key((KeyT) (((uint64_t) a.y << 32) | a.x)); count_key((KeyT) (((uint64_t) a.w << 32) | a.z));
                count_key((KeyT) (((uint64_t) b.y << 32) | b.x)); count_key((KeyT) (((uint64_t) b.w << 32) | b.z));
            }
        }
        for (; v < nvec; v += HIST_THREADS) {
            const uint4 a = ld_stream(vbase + v);
            if (sizeof(KeyT) == 4) {
                count_key((KeyT) a.x); count_key((KeyT) a.y); count_key((KeyT) a.z); count_key((KeyT) a.w);
            } else {
                count_key((KeyT) (((uint64_t) a.y << 32) | a.x)); count_key((KeyT) (((uint64_t) a.w << 32) | a.z));
            }
        }
        const uint64_t tail0 = head + nvec * VEC;
        if (tail0 + tid < count) count_key(base[tail0 + tid]);
    }
    __syncthreads();

    // Fold the 32 lane columns of every (pass, digit) and add to the global histogram.
    for (int c = tid; c < NUM_PASSES * 256; c += HIST_THREADS) {
        const int p = c >> 8, d = c & 255;
        const uint32_t *row = &hist_smem[(p * 128 + (d >> 1)) * 32];
        const uint32_t sh = (d & 1) * 16;
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += (row[(j + (d >> 1)) & 31] >> sh) & 0xffffu; // skewed: conflict-free
        if (sum) atomicAdd(&ghist[c], sum);
    }

    // Last CTA to finish turns each 256-bin histogram into its exclusive scan (the
    // cross-bin prefix of multi_radixsort.comp:64-65,74, computed once instead of per workgroup).
    __shared__ uint32_t s_last;
    __shared__ uint32_t s_scan[8];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int p = 0; p < NUM_PASSES; ++p) {
            uint32_t v = 0;
            if (tid < 256) v = __ldcg(&ghist[p * 256 + tid]);
            const uint32_t ex = block_exclusive_scan_256(v, s_scan, nullptr);
            if (tid < 256) ghist[p * 256 + tid] = ex;
            __syncthreads();
        }
        if (tid == 0) *done_counter = 0;
    }
}

// =====================================================================================
// Fused path, kernel 2: one 8-bit digit pass in a single sweep over the keys (read 4 B,
// write 4 B per key).  Tiles are handed out in ticket order; a tile learns how many keys
// of each digit precede it from the status words of earlier tiles (chained scan with
// decoupled look-back) instead of the reference's every-workgroup-reads-every-histogram
// prologue (multi_radixsort.comp:56-62, O(W^2) traffic).
//   status[tile][digit]        this pass: 0 -> AGGREGATE|count -> INCLUSIVE|prefix
//   status_clear[tile][digit]  the array the next pass will use; zeroed here so no memset
//                              sits between passes.
// =====================================================================================
struct ChainedScanBase {
    uint32_t *status, *status_clear;
    const uint32_t *bin_start;
    uint32_t *error_flag;
    uint32_t tile;

    __device__ __forceinline__ void publish(uint32_t d, uint32_t count) const {
        st_relaxed_gpu(status + (size_t) tile * RADIX + d,
                       (tile == 0 ? STATUS_FLAG_INCLUSIVE : STATUS_FLAG_AGGREGATE) | count);
        if (status_clear) status_clear[(size_t) tile * RADIX + d] = 0;
    }
    // Windowed look-back: LOOKBACK_WINDOW earlier tiles are polled together (independent loads in
    // flight), then consumed nearest-first up to the first INCLUSIVE word.  On B200 a tile retires
    // every few tens of ns while one L2 round trip is a few hundred, so ~10 predecessors have only
    // their AGGREGATE out when a tile looks back: walking them one load at a time is what stalls.
    template <int WINDOW = LOOKBACK_WINDOW, int BACKOFF_NS = 0>
    __device__ __forceinline__ uint32_t resolve(uint32_t d, uint32_t count) const {
        uint32_t excl = 0;
        if (tile != 0) {
            uint32_t *my = status + (size_t) tile * RADIX + d;
            uint32_t back = 1; // distance of the nearest tile not yet consumed
            uint32_t spins = 0;
            bool done = false;
            while (!done) {
                uint32_t w[WINDOW];
#pragma unroll
                for (int i = 0; i < WINDOW; ++i)
                    w[i] = (back + i <= tile) ? ld_relaxed_gpu(my - (size_t) (back + i) * RADIX) : 0u;
#pragma unroll
                for (int i = 0; i < WINDOW; ++i) {
                    if (done || (w[i] & STATUS_FLAG_MASK) == 0) break; // unpublished: poll again from here
                    excl += w[i] & STATUS_VALUE_MASK;
                    back++;
                    if (w[i] & STATUS_FLAG_INCLUSIVE) done = true; // tile 0 always publishes INCLUSIVE
                }
                if (++spins > LOOKBACK_SPIN_LIMIT) {
                    atomicExch(error_flag, (uint32_t) DEVERR_LOOKBACK_TIMEOUT);
                    break;
                }
                if (BACKOFF_NS > 0 && !done && (w[0] & STATUS_FLAG_MASK) == 0) __nanosleep(BACKOFF_NS); // nothing new: leave the issue slots to the workers
            }
            st_relaxed_gpu(my, STATUS_FLAG_INCLUSIVE | ((excl + count) & STATUS_VALUE_MASK));
        }
        return bin_start[d] + excl;
    }
};

// Running per-digit offsets in shared memory: the staged and single-workgroup paths, where one
// CTA walks its slab tile by tile (global_offsets[] of multi_radixsort.comp:75-76,120-122).
struct RunningOffsetsBase {
    uint32_t *global_offsets;
    __device__ __forceinline__ void publish(uint32_t, uint32_t) const {}
    __device__ __forceinline__ uint32_t resolve(uint32_t d, uint32_t count) const {
        const uint32_t g = global_offsets[d];
        global_offsets[d] = g + count;
        return g;
    }
};

template <typename KeyT, bool HAS_VALUES, int THREADS, int KPT, int MATCH, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_pass_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out,
                     const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out, uint32_t n,
                     uint32_t shift, const uint32_t *__restrict__ bin_start, uint32_t *status,
                     uint32_t *status_clear, uint32_t *ticket, uint32_t *error_flag) {
    using Sorter = TileSorter<KeyT, HAS_VALUES, THREADS, KPT, MATCH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename Sorter::Smem &s = *reinterpret_cast<typename Sorter::Smem *>(smem_raw);
    constexpr uint32_t TILE = Sorter::TILE;

    if (threadIdx.x == 0) s.misc[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s.misc[0];
    const uint64_t tile_base = (uint64_t) tile * TILE;
    if (tile_base >= n) return; // cannot happen with grid == number of tiles; kept as a guard
    const uint32_t valid = (n - tile_base < TILE) ? (uint32_t) (n - tile_base) : TILE;

    ChainedScanBase base{status, status_clear, bin_start, error_flag, tile};
    Sorter::run(s, keys_in + tile_base, keys_out, HAS_VALUES ? vals_in + tile_base : nullptr, vals_out, valid, shift,
                base);
}

// =====================================================================================
// Staged path, stage RADIX_SORT_HISTOGRAMS (multi_radixsort_histograms.comp:31-56):
// CTA w counts the digits of its slab [w*nb*256, (w+1)*nb*256) into row w of g_histograms.
// =====================================================================================
constexpr int STAGED_HIST_THREADS = 256;

__global__ void __launch_bounds__(STAGED_HIST_THREADS)
staged_histograms_kernel(const uint32_t *__restrict__ elements_in, uint32_t *__restrict__ histograms, uint32_t n,
                         uint32_t shift, uint32_t nb) {
    __shared__ uint32_t histogram[RADIX];
    const int tid = threadIdx.x;
    histogram[tid] = 0;
    __syncthreads();
    const uint64_t slab = (uint64_t) nb * 256u;
    const uint64_t lo = (uint64_t) blockIdx.x * slab;
    const uint64_t hi = (lo + slab < n) ? lo + slab : n;
    if (lo < hi) {
        const uint32_t *base = elements_in + lo;
        const uint64_t count = hi - lo;
        if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) {
            const uint64_t nvec = count / 4;
            const uint4 *vbase = reinterpret_cast<const uint4 *>(base);
            for (uint64_t v = tid; v < nvec; v += STAGED_HIST_THREADS) {
                const uint4 a = ld_stream(vbase + v);
                atomicAdd(&histogram[(a.x >> shift) & 255u], 1u);
                atomicAdd(&histogram[(a.y >> shift) & 255u], 1u);
                atomicAdd(&histogram[(a.z >> shift) & 255u], 1u);
                atomicAdd(&histogram[(a.w >> shift) & 255u], 1u);
            }
            if (nvec * 4 + tid < count) atomicAdd(&histogram[(base[nvec * 4 + tid] >> shift) & 255u], 1u);
        } else {
            for (uint64_t e = tid; e < count; e += STAGED_HIST_THREADS)
                atomicAdd(&histogram[(ld_stream(base + e) >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    histograms[(size_t) RADIX * blockIdx.x + tid] = histogram[tid];
}

// =====================================================================================
// Staged path, offsets from the histogram matrix: what every workgroup of the reference
// recomputes for itself in the prologue of multi_radixsort.comp:56-77, done once:
//   offsets[w][b] = sum_{b'<b} sum_{w'} hist[w'][b']  +  sum_{w'<w} hist[w'][b]
// as three small kernels over chunks of STAGED_CHUNK_ROWS rows.
// =====================================================================================
constexpr int STAGED_CHUNK_ROWS = 64;

__global__ void __launch_bounds__(256)
staged_colsum_kernel(const uint32_t *__restrict__ histograms, uint32_t W, uint32_t *__restrict__ chunk_sums) {
    const uint32_t c = blockIdx.x, d = threadIdx.x;
    const uint32_t r0 = c * STAGED_CHUNK_ROWS, r1 = min(W, r0 + STAGED_CHUNK_ROWS);
    uint32_t sum = 0;
    for (uint32_t r = r0; r < r1; ++r) sum += histograms[(size_t) r * RADIX + d];
    chunk_sums[(size_t) c * RADIX + d] = sum;
}

__global__ void __launch_bounds__(256)
staged_chunkscan_kernel(uint32_t *chunk_sums, uint32_t num_chunks, uint32_t *bin_start) {
    __shared__ uint32_t s_scan[8];
    const uint32_t d = threadIdx.x;
    uint32_t run = 0;
    for (uint32_t c = 0; c < num_chunks; ++c) { // exclusive prefix over chunks, in place
        const uint32_t t = chunk_sums[(size_t) c * RADIX + d];
        chunk_sums[(size_t) c * RADIX + d] = run;
        run += t;
    }
    bin_start[d] = block_exclusive_scan_256(run, s_scan, nullptr);
}

__global__ void __launch_bounds__(256)
staged_offsets_kernel(const uint32_t *__restrict__ histograms, uint32_t W, const uint32_t *__restrict__ chunk_sums,
                      const uint32_t *__restrict__ bin_start, uint32_t *__restrict__ offsets) {
    const uint32_t c = blockIdx.x, d = threadIdx.x;
    const uint32_t r0 = c * STAGED_CHUNK_ROWS, r1 = min(W, r0 + STAGED_CHUNK_ROWS);
    uint32_t run = bin_start[d] + chunk_sums[(size_t) c * RADIX + d];
    for (uint32_t r = r0; r < r1; ++r) {
        offsets[(size_t) r * RADIX + d] = run;
        run += histograms[(size_t) r * RADIX + d];
    }
}

// =====================================================================================
// Staged path, stage RADIX_SORT body (multi_radixsort.comp:80-126): CTA w stably scatters
// its slab, sub-tile by sub-tile, from running per-digit offsets kept in shared memory
// (the role of global_offsets[], bumped per block at :120-122).
// =====================================================================================
template <bool HAS_VALUES, int THREADS, int KPT, int MATCH, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
staged_scatter_kernel(const uint32_t *__restrict__ elements_in, uint32_t *__restrict__ elements_out,
                      const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out,
                      const uint32_t *__restrict__ offsets, uint32_t n, uint32_t shift, uint32_t nb) {
    using Sorter = TileSorter<uint32_t, HAS_VALUES, THREADS, KPT, MATCH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename Sorter::Smem &s = *reinterpret_cast<typename Sorter::Smem *>(smem_raw);
    __shared__ uint32_t global_offsets[RADIX];
    constexpr uint32_t TILE = Sorter::TILE;

    if (threadIdx.x < RADIX) global_offsets[threadIdx.x] = offsets[(size_t) blockIdx.x * RADIX + threadIdx.x];
    const uint64_t slab = (uint64_t) nb * 256u;
    const uint64_t lo = (uint64_t) blockIdx.x * slab;
    const uint64_t hi = (lo + slab < n) ? lo + slab : n;
    RunningOffsetsBase base_fn{global_offsets};
    __syncthreads();
    for (uint64_t t0 = lo; t0 < hi; t0 += TILE) {
        const uint32_t valid = (hi - t0 < TILE) ? (uint32_t) (hi - t0) : TILE;
        Sorter::run(s, elements_in + t0, elements_out, HAS_VALUES ? vals_in + t0 : nullptr, vals_out, valid, shift,
                    base_fn);
        __syncthreads();
    }
}

// =====================================================================================
// Single-workgroup path (single_radixsort.comp:42-139): one CTA does all four passes,
// ping-ponging buf0 -> buf1 -> buf0 -> buf1 -> buf0 inside the kernel; the result ends in
// buf0.  Pointers are deliberately not __restrict__/read-only: both buffers are read and
// written by this one CTA, ordered by block barriers.
// =====================================================================================
template <int THREADS, int KPT, int MATCH>
__global__ void __launch_bounds__(THREADS, 1)
single_sort_kernel(uint32_t *buf0, uint32_t *buf1, uint32_t n, const uint32_t *__restrict__ gate /* may be NULL: the kernel only works if *gate != 0 */) {
    grid_dependency_wait();
    if (gate && *gate == 0) return;
    using Sorter = TileSorter<uint32_t, false, THREADS, KPT, MATCH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename Sorter::Smem &s = *reinterpret_cast<typename Sorter::Smem *>(smem_raw);
    __shared__ uint32_t histogram[RADIX];
    __shared__ uint32_t global_offsets[RADIX];
    __shared__ uint32_t s_scan[8];
    constexpr uint32_t TILE = Sorter::TILE;
    const int tid = threadIdx.x;

    for (uint32_t iteration = 0; iteration < 4; ++iteration) { // ITERATIONS, single_radixsort.comp:14
        const uint32_t shift = 8 * iteration;
        uint32_t *src = (iteration & 1) ? buf1 : buf0; // ELEMENT_IN, :40
        uint32_t *dst = (iteration & 1) ? buf0 : buf1; // :129-133
        if (tid < RADIX) histogram[tid] = 0;
        __syncthreads();
        for (uint32_t e = tid; e < n; e += THREADS) atomicAdd(&histogram[(src[e] >> shift) & 255u], 1u); // :56-61
        __syncthreads();
        const uint32_t ex = block_exclusive_scan_256(tid < RADIX ? histogram[tid] : 0u, s_scan, nullptr); // :65-84
        if (tid < RADIX) global_offsets[tid] = ex;
        __syncthreads();
        RunningOffsetsBase base_fn{global_offsets};
        for (uint32_t t0 = 0; t0 < n; t0 += TILE) { // :91
            const uint32_t valid = (n - t0 < TILE) ? (n - t0) : TILE;
            Sorter::run(s, src + t0, dst, nullptr, nullptr, valid, shift, base_fn);
            __syncthreads();
        }
        __threadfence_block();
        __syncthreads();
    }
}

} // namespace vkrs
