// vkrs_api.cu -- host side of the C-ABI declared in include/vkradixsort_b200.h.
//
// This file owns launch configuration, the handle's workspace and argument checking; it
// contains no sorting logic and no CPU fallback: every entry point either enqueues the
// sm_100a kernels of vkrs_kernels.cuh or returns an error.
#include "../../include/vkradixsort_b200.h"
#include "vkrs_kernels.cuh"
#include "vkrs_msd.cuh"
#include "vkrs_exchange.cuh"
#include "vkrs_pipeline.cuh"
#include "vkrs_segmented.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

using namespace vkrs;

// ---- whole-sort schedules / tile configurations (selectable for tuning runs) --------------
struct PassConfig {
    const char *name;
    int threads, kpt, min_blocks;
};

constexpr int NUM_VARIANTS = 8;
// Keep in sync with launch_pass_u32() below.
//   "seg"    = segment_histogram_kernel + segmented_scatter_kernel (count first, no inter-CTA chain)
//   "pipe"   = onesweep_pipelined_kernel (single sweep, persistent, look-back on a control group)
//   "simple" = onesweep_pass_kernel (single sweep, one tile per CTA)
const PassConfig kVariants[NUM_VARIANTS] = {
    {"seg 2x384x16", 384, 16, 1},      {"seg 2x384x20", 384, 20, 1},       {"seg 2x352x18", 352, 18, 1},
    {"seg 1x512x12 2cta", 512, 12, 2}, {"pipe 512x16 1cta", 512, 16, 1},   {"pipe 256x24 r96/32", 256, 24, 2},
    {"simple 512x16 ptx", 512, 16, 2}, {"simple 512x16 ballot", 512, 16, 2},
};
constexpr int DEFAULT_VARIANT = 0;

// default configurations of the other paths
constexpr int PAIR_WORKERS = 352, PAIR_KPT = 16; // segmented path, two worker groups per CTA
constexpr int U64_WORKERS = 256, U64_KPT = 16;
constexpr int STAGED_THREADS = 256, STAGED_KPT = 16;
constexpr int SINGLE_THREADS = 1024, SINGLE_KPT = 8;
constexpr uint32_t AUTO_SINGLE_MAX = 12288; // vkrs_sort_auto: single path up to here (measured crossover ~1.5e4, profiles/r01_nsweep.jsonl)

thread_local std::string g_create_error;

} // namespace

struct vkrs_context {
    int device = 0;
    int sm_count = 148;
    int variant = DEFAULT_VARIANT;
    uint64_t launches = 0;
    std::string error;

    // control block: [8][256] global histograms | 8 tickets | done counter | error flag
    uint32_t *ctrl = nullptr;
    static constexpr size_t CTRL_HIST = 8 * 256;
    static constexpr size_t CTRL_TICKETS = CTRL_HIST;
    static constexpr size_t CTRL_DONE = CTRL_HIST + 8;
    static constexpr size_t CTRL_ERROR = CTRL_HIST + 9;
    static constexpr size_t CTRL_WORDS = CTRL_HIST + 16;

    // chained-scan tile status, two arrays used alternately by consecutive passes
    uint32_t *status[2] = {nullptr, nullptr};
    uint64_t status_rows = 0;

    // segmented path: hist[segments][256]
    uint32_t *seg_hist = nullptr;
    uint64_t seg_hist_rows = 0;

    // staged path: offsets[W][256], chunk sums, bin starts
    uint32_t *staged_offsets = nullptr;
    uint64_t staged_rows = 0;
    uint32_t *staged_chunks = nullptr;
    uint64_t staged_chunk_rows = 0;
    uint32_t *staged_bin_start = nullptr;

    // opt-in per-kernel timing (vkrs_set_profiling): one event pair per launch
    bool profiling = false;
    struct ProfRecord {
        const char *name;
        cudaEvent_t start, stop;
    };
    std::vector<ProfRecord> prof_records;
    std::vector<std::pair<std::string, std::pair<double, uint64_t>>> prof_summary; // name -> (ms, launches)

    // phase timers of the pipelined kernel (tuning aid; NULL unless vkrs_debug_counters was enabled)
    unsigned long long *debug_counters = nullptr;

    // vkrs_multi_sort_host device buffers; events around its three stages and what they measured last (ms)
    uint32_t *host_buf[2] = {nullptr, nullptr};
    uint64_t host_cap = 0;
    cudaEvent_t host_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double host_ms[3] = {0, 0, 0};

    // keys-only whole-sort schedule (vkrs_set_schedule) and the bucket schedule's workspace (vkrs_msd.cuh)
    int schedule = 0; // VKRS_SCHEDULE_AUTO
    unsigned char *msd_ws = nullptr;
    uint32_t msd_segments_cap = 0; // segments the workspace was laid out for
    uint32_t msd_plan_n = 0, msd_plan_segments = 0, msd_plan_seg_keys = 0; // cached pass-1 piece table
    int msd_stop_after = 0; // vkrs_debug_bucket_stop: 0 = run the whole schedule
    bool msd_guess_window = true; // VKRS_GUESS_WINDOW=0 turns the sampled guess of the digit window off (tests of the recount path)
    uint32_t msd_first_shift = 24, msd_first_base = 0; // digit window the first histogram of the bucket schedule counts in (vkrs_set_key_span_hint)
    uint32_t msd_use_bins = 1; // local sort: 0 = per-bucket path only (VKRS_LOCAL_BINS=0, tests)
    uint32_t *msd_items = nullptr; // item_first[items + 1] | item_lo[items + 1] of the local sort
    uint64_t msd_items_cap = 0;
};

namespace {

int fail(vkrs_context *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->error = buf;
    else g_create_error = buf;
    return code;
}

 // Brackets one kernel launch with an event pair when profiling is on; counts the launch.
struct LaunchScope {
    vkrs_context *h;
    cudaStream_t stream;
    cudaEvent_t stop = nullptr;
    LaunchScope(vkrs_context *h_, const char *name, cudaStream_t s) : h(h_), stream(s) {
        h->launches++;
        if (!h->profiling) return;
        cudaEvent_t start = nullptr;
        if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) {
            stop = nullptr;
            return;
        }
        cudaEventRecord(start, stream);
        h->prof_records.push_back({name, start, stop});
    }
    ~LaunchScope() {
        if (stop) cudaEventRecord(stop, stream);
    }
};

#define VKRS_CUDA(h, expr)                                                                                     \
    do {                                                                                                       \
        cudaError_t e__ = (expr);                                                                              \
        if (e__ != cudaSuccess)                                                                                \
            return fail(h, VKRS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,   \
                        __LINE__);                                                                             \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            active = true;
        }
    }
    ~DeviceGuard() {
        if (active) cudaSetDevice(prev);
    }
};

template <typename T>
int grow(vkrs_context *h, T *&ptr, uint64_t &cap, uint64_t want, size_t elem_bytes, bool zero) {
    if (want <= cap && ptr) return VKRS_OK;
    // Growing is the one place a call may synchronise: the old workspace can still be in use.
    VKRS_CUDA(h, cudaDeviceSynchronize());
    if (ptr) VKRS_CUDA(h, cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    void *p = nullptr;
    VKRS_CUDA(h, cudaMalloc(&p, want * elem_bytes));
    if (zero) VKRS_CUDA(h, cudaMemset(p, 0, want * elem_bytes));
    ptr = static_cast<T *>(p);
    cap = want;
    return VKRS_OK;
}

int ensure_status(vkrs_context *h, uint64_t tiles) {
    if (tiles <= h->status_rows) return VKRS_OK;
    const uint64_t want = tiles + tiles / 8 + 64;
    uint64_t c0 = h->status_rows, c1 = h->status_rows;
    int r = grow(h, h->status[0], c0, want * RADIX, sizeof(uint32_t), true);
    if (r) return r;
    r = grow(h, h->status[1], c1, want * RADIX, sizeof(uint32_t), true);
    if (r) return r;
    h->status_rows = want;
    return VKRS_OK;
}

// Launch with programmatic dependent launch enabled: the grid may be scheduled while the previous
// kernel of the stream is finishing; the kernels call grid_dependency_wait() before consuming.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename Kernel>
int set_smem(vkrs_context *h, Kernel kernel, size_t bytes) {
    VKRS_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    return VKRS_OK;
}

template <typename KeyT, bool HAS_VALUES, int THREADS, int KPT, int MATCH, int MIN_BLOCKS>
int launch_pass_t(vkrs_context *h, const KeyT *in, KeyT *out, const uint32_t *vin, uint32_t *vout, uint32_t n,
                  uint32_t shift, int pass_index, cudaStream_t stream) {
    using Sorter = TileSorter<KeyT, HAS_VALUES, THREADS, KPT, MATCH>;
    auto kernel = onesweep_pass_kernel<KeyT, HAS_VALUES, THREADS, KPT, MATCH, MIN_BLOCKS>;
    static thread_local int configured_device = -1;
    if (configured_device != h->device) {
        int r = set_smem(h, kernel, sizeof(typename Sorter::Smem));
        if (r) return r;
        configured_device = h->device;
    }
    const uint32_t tiles = (uint32_t) (((uint64_t) n + Sorter::TILE - 1) / Sorter::TILE);
    {
        LaunchScope scope(h, HAS_VALUES ? "onesweep_pass_kernel<pairs>" : (sizeof(KeyT) == 8 ? "onesweep_pass_kernel<u64>" : "onesweep_pass_kernel"), stream);
        kernel<<<tiles, THREADS, sizeof(typename Sorter::Smem), stream>>>(
            in, out, vin, vout, n, shift, h->ctrl + pass_index * RADIX, h->status[pass_index & 1],
            h->status[(pass_index + 1) & 1], h->ctrl + vkrs_context::CTRL_TICKETS + pass_index,
            h->ctrl + vkrs_context::CTRL_ERROR);
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

// The pipelined kernel is persistent: the grid is the number of CTAs that are co-resident
// (every CTA must be running for the chained scan to make progress), capped by the tile count.
template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT, int MIN_BLOCKS, int REG_WORKER = 0, int REG_CTRL = 0,
          int MATCH = MATCH_TABLE, int GROUPS = 1>
int launch_pipe_t(vkrs_context *h, const KeyT *in, KeyT *out, const uint32_t *vin, uint32_t *vout, uint32_t n,
                  uint32_t shift, int pass_index, cudaStream_t stream) {
    using Smem = PipeSmem<KeyT, HAS_VALUES, WORKERS, KPT, GROUPS>;
    auto kernel = onesweep_pipelined_kernel<KeyT, HAS_VALUES, WORKERS, KPT, MIN_BLOCKS, REG_WORKER, REG_CTRL, MATCH, GROUPS>;
    static thread_local int configured_device = -1;
    static thread_local int blocks_per_sm = 0;
    if (configured_device != h->device) {
        int r = set_smem(h, kernel, sizeof(Smem));
        if (r) return r;
        VKRS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, GROUPS * WORKERS + CTRL_THREADS, sizeof(Smem)));
        if (blocks_per_sm < 1) return fail(h, VKRS_ERR_INTERNAL, "pipelined kernel does not fit on an SM");
        configured_device = h->device;
    }
    const uint64_t tiles = ((uint64_t) n + Smem::TILE - 1) / Smem::TILE;
    uint64_t grid = (uint64_t) h->sm_count * blocks_per_sm;
    if (grid > tiles) grid = tiles;
    {
        LaunchScope scope(h, HAS_VALUES ? "onesweep_pipelined_kernel<pairs>" : (sizeof(KeyT) == 8 ? "onesweep_pipelined_kernel<u64>" : "onesweep_pipelined_kernel"), stream);
        kernel<<<(unsigned) grid, GROUPS * WORKERS + CTRL_THREADS, sizeof(Smem), stream>>>(
            in, out, vin, vout, n, shift, h->ctrl + pass_index * RADIX, h->status[pass_index & 1],
            h->status[(pass_index + 1) & 1], h->ctrl + vkrs_context::CTRL_TICKETS + pass_index,
            h->ctrl + vkrs_context::CTRL_ERROR, h->debug_counters);
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

// One digit pass of the segmented path: count (one CTA per segment), then the persistent scatter.
// do_count / do_scatter let the multi-GPU path run the two halves separately (the exchange plan is
// made from the counts in between); dst_tables != nullptr selects the peer-to-peer write-out.
template <typename KeyT, bool HAS_VALUES, int WORKERS, int KPT, int GROUPS, int MIN_BLOCKS, bool PARTITION = false, bool P2P = false,
          int XF_IN = 0, int XF_OUT = 0>
int launch_seg_t(vkrs_context *h, const KeyT *in, KeyT *out, const uint32_t *vin, uint32_t *vout, uint32_t n,
                 uint32_t shift, cudaStream_t stream, uint32_t key_base = 0, uint32_t *bucket_totals = nullptr,
                 bool do_count = true, bool do_scatter = true, const unsigned long long *dst_tables = nullptr,
                 const uint32_t *gate = nullptr) {
    using Smem = SegSmem<KeyT, HAS_VALUES, WORKERS, KPT, GROUPS>;
    constexpr uint32_t TILE = Smem::Group::TILE;
    auto kernel = segmented_scatter_kernel<KeyT, HAS_VALUES, WORKERS, KPT, GROUPS, MIN_BLOCKS, PARTITION, P2P, XF_IN, XF_OUT>;
    static thread_local int configured_device = -1;
    static thread_local int blocks_per_sm = 0;
    if (configured_device != h->device) {
        int r = set_smem(h, kernel, sizeof(Smem));
        if (r) return r;
        VKRS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, GROUPS * WORKERS + 32, sizeof(Smem)));
        if (blocks_per_sm < 1) return fail(h, VKRS_ERR_INTERNAL, "segmented scatter kernel does not fit on an SM");
        configured_device = h->device;
    }
    const uint32_t tiles = (uint32_t) (((uint64_t) n + TILE - 1) / TILE);
    uint32_t ctas = (uint32_t) (h->sm_count * blocks_per_sm);
    const uint32_t need = (tiles + GROUPS - 1) / GROUPS;
    if (ctas > need) ctas = need;
    if (ctas == 0) ctas = 1;
    const uint32_t segments = ctas * GROUPS;
    int r = grow(h, h->seg_hist, h->seg_hist_rows, (uint64_t) segments, RADIX * sizeof(uint32_t), false);
    if (r) return r;
    if (do_count) {
        {
            LaunchScope scope(h, "segment_histogram_kernel", stream);
            VKRS_CUDA(h, launch_pdl(segment_histogram_kernel<KeyT, PARTITION, XF_IN>, dim3(segments), dim3(SEGHIST_THREADS), 0, stream, in, n, shift,
                                    key_base, TILE, tiles, h->seg_hist, gate));
        }
        if (bucket_totals) {
            LaunchScope scope(h, "segment_column_sum_kernel", stream);
            segment_column_sum_kernel<<<1, RADIX, 0, stream>>>(h->seg_hist, segments, bucket_totals);
        }
    }
    if (do_scatter) {
        LaunchScope scope(h, P2P ? "segmented_scatter_kernel<p2p>" : HAS_VALUES ? "segmented_scatter_kernel<pairs>" : (sizeof(KeyT) == 8 ? "segmented_scatter_kernel<u64>" : "segmented_scatter_kernel"), stream);
        VKRS_CUDA(h, launch_pdl(kernel, dim3(ctas), dim3(GROUPS * WORKERS + 32), sizeof(Smem), stream, in, out, vin, vout, n, shift,
                                key_base, (const uint32_t *) h->seg_hist, tiles, h->debug_counters, dst_tables, gate));
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

// Typed keys: the order-preserving transform is applied as pass 0 reads the keys and undone as the
// last pass writes them -- no extra pass, no extra traffic.
template <int XF>
int typed_sort_u32(vkrs_context *h, uint32_t *buf0, uint32_t *buf1, uint32_t n, cudaStream_t s) {
    int r = launch_seg_t<uint32_t, false, 384, 16, 2, 1, false, false, XF, 0>(h, buf0, buf1, nullptr, nullptr, n, 0, s);
    if (r) return r;
    r = launch_seg_t<uint32_t, false, 384, 16, 2, 1>(h, buf1, buf0, nullptr, nullptr, n, 8, s);
    if (r) return r;
    r = launch_seg_t<uint32_t, false, 384, 16, 2, 1>(h, buf0, buf1, nullptr, nullptr, n, 16, s);
    if (r) return r;
    return launch_seg_t<uint32_t, false, 384, 16, 2, 1, false, false, 0, XF>(h, buf1, buf0, nullptr, nullptr, n, 24, s);
}

int launch_pass_u32(vkrs_context *h, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t shift, int pass_index,
                    cudaStream_t stream) {
    switch (h->variant) {
        case 0: return launch_seg_t<uint32_t, false, 384, 16, 2, 1>(h, in, out, nullptr, nullptr, n, shift, stream);
        case 1: return launch_seg_t<uint32_t, false, 384, 20, 2, 1>(h, in, out, nullptr, nullptr, n, shift, stream);
        case 2: return launch_seg_t<uint32_t, false, 352, 18, 2, 1>(h, in, out, nullptr, nullptr, n, shift, stream);
        case 3: return launch_seg_t<uint32_t, false, 512, 12, 1, 2>(h, in, out, nullptr, nullptr, n, shift, stream);
        case 4: return launch_pipe_t<uint32_t, false, 512, 16, 1>(h, in, out, nullptr, nullptr, n, shift, pass_index, stream);
        case 5: return launch_pipe_t<uint32_t, false, 256, 24, 2, 96, 32>(h, in, out, nullptr, nullptr, n, shift, pass_index, stream);
        case 6: return launch_pass_t<uint32_t, false, 512, 16, MATCH_PTX, 2>(h, in, out, nullptr, nullptr, n, shift, pass_index, stream);
        case 7: return launch_pass_t<uint32_t, false, 512, 16, MATCH_BALLOT, 2>(h, in, out, nullptr, nullptr, n, shift, pass_index, stream);
        default: return fail(h, VKRS_ERR_INVALID_ARGUMENT, "unknown kernel variant %d", h->variant);
    }
}

bool variant_is_segmented(int v) { return strncmp(kVariants[v].name, "seg", 3) == 0; }

uint32_t variant_tile(int v) { return (uint32_t) (kVariants[v].threads * kVariants[v].kpt); }

// Zero the control block and build the exclusive-scanned global histograms.
template <typename KeyT, int NUM_PASSES>
int launch_global_histogram(vkrs_context *h, const KeyT *keys, uint32_t n, cudaStream_t stream) {
    auto kernel = global_histogram_kernel<KeyT, NUM_PASSES>;
    const size_t smem = (size_t) NUM_PASSES * 128 * 32 * sizeof(uint32_t);
    static thread_local int configured_device = -1;
    if (configured_device != h->device) {
        int r = set_smem(h, kernel, smem);
        if (r) return r;
        configured_device = h->device;
    }
    VKRS_CUDA(h, cudaMemsetAsync(h->ctrl, 0, (vkrs_context::CTRL_DONE + 1) * sizeof(uint32_t), stream));
    const int ctas_per_sm = (NUM_PASSES <= 4) ? 3 : 1;
    uint64_t grid = (uint64_t) h->sm_count * ctas_per_sm;
    const uint64_t min_grid = ((uint64_t) n + HIST_MAX_KEYS_PER_CTA - 1) / HIST_MAX_KEYS_PER_CTA;
    if (grid < min_grid) grid = min_grid;
    // per-CTA chunk: multiple of 2048 keys so that every CTA but the last starts 16-byte aligned
    uint64_t per_cta = ((uint64_t) n + grid - 1) / grid;
    per_cta = (per_cta + 2047) / 2048 * 2048;
    if (per_cta == 0) per_cta = 2048;
    grid = ((uint64_t) n + per_cta - 1) / per_cta;
    if (grid == 0) grid = 1;
    {
        LaunchScope scope(h, "global_histogram_kernel", stream);
        kernel<<<(unsigned) grid, HIST_THREADS, smem, stream>>>(keys, n, (uint32_t) per_cta, h->ctrl,
                                                                 h->ctrl + vkrs_context::CTRL_DONE);
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

// ---- the bucket schedule (vkrs_msd.cuh) -------------------------------------------------------
#ifndef VKRS_MSD_WORKERS
#define VKRS_MSD_WORKERS 384
#define VKRS_MSD_KPT 20
#define VKRS_MSD_GROUPS 2
#endif
#ifndef VKRS_MSD_UNIFORM_FAST
#define VKRS_MSD_UNIFORM_FAST 1
#endif
constexpr int MSD_WORKERS = VKRS_MSD_WORKERS, MSD_KPT = VKRS_MSD_KPT, MSD_GROUPS = VKRS_MSD_GROUPS;
constexpr bool MSD_UNIFORM_FAST = VKRS_MSD_UNIFORM_FAST != 0;
constexpr uint32_t MSD_TILE = MSD_WORKERS * MSD_KPT;
constexpr uint32_t MSD_SUBS = RADIX * RADIX;          // (digit1, digit2) buckets
// vkrs_multi_sort, schedule auto (measured crossovers, profiles/r01_schedule_sweep.jsonl): the bucket schedule wins
// from 4*10^6 keys up; above 2.2*10^8 uniform keys a 16-bit-prefix bucket passes LOCAL_MAX keys and the schedule would
// fall back, so the LSD passes are chosen directly -- with an unstable first pass from 3.2*10^7 keys up.
constexpr uint32_t AUTO_BUCKET_MIN = 1u << 22;
constexpr uint32_t AUTO_BUCKET_MAX = 220000000u;
constexpr uint32_t AUTO_UNSTABLE_FIRST_MIN = 1u << 25;

// Workspace of the bucket schedule, laid out for `segments` segments.
struct MsdWorkspace {
    MsdPlan *plan;
    uint4 *pieces[2];
    uint32_t *seg_first[2], *bucket_first[2];
    uint32_t *bucket_start; // 257: starts of the 256 buckets of pass 1
    uint32_t *sub_start;    // 65537: starts of the (digit1, digit2) buckets
    uint32_t *hist[2];
    BigBucket *big;         // buckets above LOCAL_MAX keys (counting path)
    uint32_t *big_done;     // finished histogram chunks per big bucket (zero between sorts)
    uint32_t *big_vcs;      // first positions of every BIG_VCHUNK values, per big bucket
    uint32_t *big_pool;     // the counters of all big buckets (zero between sorts)
    size_t bytes;
    MsdWorkspace(unsigned char *base, uint32_t segments) {
        const size_t max_pieces = (size_t) segments + RADIX;
        size_t off = 0;
        auto take = [&](size_t b) {
            const size_t at = off;
            off += (b + 255) / 256 * 256;
            return base + at;
        };
        plan = reinterpret_cast<MsdPlan *>(take(sizeof(MsdPlan)));
        for (int p = 0; p < 2; ++p) {
            pieces[p] = reinterpret_cast<uint4 *>(take(max_pieces * sizeof(uint4)));
            seg_first[p] = reinterpret_cast<uint32_t *>(take(((size_t) segments + 1) * sizeof(uint32_t)));
            bucket_first[p] = reinterpret_cast<uint32_t *>(take((RADIX + 1) * sizeof(uint32_t)));
            hist[p] = reinterpret_cast<uint32_t *>(take(max_pieces * RADIX * sizeof(uint32_t)));
        }
        bucket_start = reinterpret_cast<uint32_t *>(take((RADIX + 1) * sizeof(uint32_t)));
        sub_start = reinterpret_cast<uint32_t *>(take(((size_t) MSD_SUBS + 1) * sizeof(uint32_t)));
        big = reinterpret_cast<BigBucket *>(take((size_t) BIG_MAX * sizeof(BigBucket)));
        big_done = reinterpret_cast<uint32_t *>(take((size_t) BIG_MAX * sizeof(uint32_t)));
        big_vcs = reinterpret_cast<uint32_t *>(take((size_t) BIG_MAX * BIG_VCS_STRIDE * sizeof(uint32_t)));
        big_pool = reinterpret_cast<uint32_t *>(take((size_t) BIG_POOL_WORDS * sizeof(uint32_t)));
        bytes = off;
    }
};

using MsdScatterSmem = MsdSmem<MSD_WORKERS, MSD_KPT, MSD_GROUPS>;

int msd_prepare(vkrs_context *h, uint32_t n, uint32_t &ctas, uint32_t &segments, uint32_t &seg_keys) {
    auto kernel = msd_scatter_kernel<MSD_WORKERS, MSD_KPT, MSD_GROUPS, MSD_UNIFORM_FAST>;
    static thread_local int configured_device = -1;
    static thread_local int blocks_per_sm = 0;
    if (configured_device != h->device) {
        int r = set_smem(h, kernel, sizeof(MsdScatterSmem));
        if (r) return r;
        r = set_smem(h, msd_local_tile_kernel<0>, sizeof(LocalTileSmem));
        if (r) return r;
        VKRS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, MSD_GROUPS * MSD_WORKERS + 32, sizeof(MsdScatterSmem)));
        if (blocks_per_sm < 1) return fail(h, VKRS_ERR_INTERNAL, "bucket scatter kernel does not fit on an SM");
        configured_device = h->device;
    }
    const uint32_t tiles = (uint32_t) (((uint64_t) n + MSD_TILE - 1) / MSD_TILE);
    ctas = (uint32_t) (h->sm_count * blocks_per_sm);
    const uint32_t need = (tiles + MSD_GROUPS - 1) / MSD_GROUPS;
    if (ctas > need) ctas = need;
    if (ctas == 0) ctas = 1;
    segments = ctas * MSD_GROUPS;
    if (segments > (uint32_t) MSD_PLAN_THREADS) return fail(h, VKRS_ERR_INTERNAL, "%u segments: the piece planner takes at most %d", segments, MSD_PLAN_THREADS);
    // a whole number of tiles per segment: every full tile of pass 1 is 16-byte aligned for TMA
    const uint32_t tiles_per_seg = (tiles + segments - 1) / segments;
    seg_keys = tiles_per_seg * MSD_TILE;
    if (!h->msd_ws || segments > h->msd_segments_cap) {
        const uint32_t cap = (uint32_t) (h->sm_count * blocks_per_sm * MSD_GROUPS);
        VKRS_CUDA(h, cudaDeviceSynchronize());
        if (h->msd_ws) VKRS_CUDA(h, cudaFree(h->msd_ws));
        h->msd_ws = nullptr;
        h->msd_plan_n = 0;
        const size_t bytes = MsdWorkspace(nullptr, cap).bytes;
        void *p = nullptr;
        VKRS_CUDA(h, cudaMalloc(&p, bytes));
        VKRS_CUDA(h, cudaMemset(p, 0, bytes));
        h->msd_ws = static_cast<unsigned char *>(p);
        h->msd_segments_cap = cap;
    }
    return VKRS_OK;
}

// One digit pass of the bucket machinery: histogram of every piece, then the unstable scatter.  with_or: the
// histogram also gathers the smallest and the largest key; gate: the histogram only works if *gate != 0 (recount);
// do_count / do_scatter select the halves.
template <int XF = 0>
int msd_pass(vkrs_context *h, const MsdWorkspace &w, int pass, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t ctas,
             uint32_t num_count_ctas, const uint32_t *bucket_start, uint32_t *sub_start, uint32_t max_sub, bool with_or,
             const uint32_t *gate, bool do_count, bool do_scatter, cudaStream_t s) {
    const uint4 *pieces = w.pieces[pass];
    const uint32_t *num_pieces = &w.plan->num_pieces[pass];
    if (do_count) {
        LaunchScope scope(h, gate ? "msd_piece_histogram_kernel<recount>" : "msd_piece_histogram_kernel", s);
        if (with_or)
            VKRS_CUDA(h, launch_pdl(msd_piece_histogram_kernel<true, XF>, dim3(num_count_ctas), dim3(MSD_HIST_THREADS), 0, s, in, pieces, num_pieces,
                                    w.plan, pass, w.hist[pass], gate));
        else
            VKRS_CUDA(h, launch_pdl(msd_piece_histogram_kernel<false, XF>, dim3(num_count_ctas), dim3(MSD_HIST_THREADS), 0, s, in, pieces, num_pieces,
                                    w.plan, pass, w.hist[pass], gate));
    }
    if (do_scatter) {
        LaunchScope scope(h, "msd_scatter_kernel", s);
        auto kernel = msd_scatter_kernel<MSD_WORKERS, MSD_KPT, MSD_GROUPS, MSD_UNIFORM_FAST, XF>;
        if (XF != 0) {
            static thread_local int configured_device = -1;
            if (configured_device != h->device) {
                int r = set_smem(h, kernel, sizeof(MsdScatterSmem));
                if (r) return r;
                configured_device = h->device;
            }
        }
        VKRS_CUDA(h, launch_pdl(kernel, dim3(ctas), dim3(MSD_GROUPS * MSD_WORKERS + 32),
                                sizeof(MsdScatterSmem), s, in, out, n, w.plan, pass, pieces, (const uint32_t *) w.seg_first[pass],
                                (const uint32_t *) w.bucket_first[pass], bucket_start, (const uint32_t *) w.hist[pass], sub_start, max_sub));
    }
    return VKRS_OK;
}

// init + (cached per N) the piece table of the first pass: pieces == segments, one bucket [0, n)
// (guess_keys != NULL: no key-span hint was given -- the init kernel guesses the digit window from a sample of these keys)
template <int XF = 0>
int msd_begin(vkrs_context *h, const MsdWorkspace &w, uint32_t n, uint32_t segments, uint32_t seg_keys, uint32_t shift0,
              uint32_t shift1, uint32_t base0, cudaStream_t s, const uint32_t *guess_keys = nullptr) {
    {
        LaunchScope scope(h, "msd_init_kernel", s);
        VKRS_CUDA(h, launch_pdl(msd_init_kernel<XF>, dim3(1), dim3(MSD_GUESS_THREADS), 0, s, w.plan, shift0, shift1, base0, guess_keys, n,
                                guess_keys != nullptr ? 1u : 0u));
    }
    if (h->msd_plan_n != n || h->msd_plan_segments != segments || h->msd_plan_seg_keys != seg_keys) {
        LaunchScope scope(h, "msd_plan_pieces_kernel", s);
        VKRS_CUDA(h, launch_pdl(msd_plan_pieces_kernel, dim3(1), dim3(MSD_PLAN_THREADS), 0, s, (const uint32_t *) nullptr, 1u, n, seg_keys,
                                segments, w.pieces[0], w.seg_first[0], w.bucket_first[0], &w.plan->num_pieces[0], (uint32_t *) nullptr,
                                (const uint32_t *) nullptr));
        h->msd_plan_n = n;
        h->msd_plan_segments = segments;
        h->msd_plan_seg_keys = seg_keys;
    }
    return VKRS_OK;
}

// Keys-only sort, least significant digit first, whose FIRST pass is the unstable scatter: a first
// pass has no earlier order to keep.  buf0 -> buf1 -> buf0 -> buf1 -> buf0 as MultiRadixSort.cpp:34-62.
int lsd_unstable_first_sort_u32(vkrs_context *h, uint32_t *buf0, uint32_t *buf1, uint32_t n, cudaStream_t s) {
    uint32_t ctas, segments, seg_keys;
    int r = msd_prepare(h, n, ctas, segments, seg_keys);
    if (r) return r;
    const MsdWorkspace w(h->msd_ws, h->msd_segments_cap);
    r = msd_begin(h, w, n, segments, seg_keys, 0, 0, 0, s);
    if (r) return r;
    r = msd_pass(h, w, 0, buf0, buf1, n, ctas, segments, nullptr, nullptr, 0, false, nullptr, true, true, s);
    if (r) return r;
    for (int p = 1; p < 4; ++p) {
        uint32_t *in = (p & 1) ? buf1 : buf0, *out = (p & 1) ? buf0 : buf1;
        r = launch_seg_t<uint32_t, false, 384, 16, 2, 1>(h, in, out, nullptr, nullptr, n, 8 * p, s);
        if (r) return r;
    }
    return VKRS_OK;
}

// The bucket schedule (see vkrs_msd.cuh).  Result in buf0.  XF != 0 (typed keys, KeyXform): pass 1 reads the keys
// through the order-preserving map, everything in between works on mapped keys, the local sort (or the last
// fallback pass) writes them back through the inverse map -- no extra pass, no extra traffic.
template <int XF>
int msd_sort_t(vkrs_context *h, uint32_t *buf0, uint32_t *buf1, uint32_t n, cudaStream_t s) {
    uint32_t ctas, segments, seg_keys;
    int r = msd_prepare(h, n, ctas, segments, seg_keys);
    if (r) return r;
    const MsdWorkspace w(h->msd_ws, h->msd_segments_cap);
    const uint32_t shift0 = XF == 0 ? h->msd_first_shift : 24u; // the key-span hint is about raw keys
    // no key-span hint: the digit window is guessed from a sample, so that keys that do not fill the 32 bits are not counted twice
    const bool guess = h->msd_guess_window && (XF != 0 || (h->msd_first_shift == 24 && h->msd_first_base == 0));
    r = msd_begin<XF>(h, w, n, segments, seg_keys, shift0, shift0 - 8, XF == 0 ? h->msd_first_base : 0u, s, guess ? (const uint32_t *) buf0 : nullptr);
    if (r) return r;
    // ---- pass 1: top digit, whole array = one bucket ----
    r = msd_pass<XF>(h, w, 0, buf0, buf1, n, ctas, segments, nullptr, nullptr, 0, true, nullptr, true, false, s);
    if (r) return r;
    {
        LaunchScope scope(h, "msd_window_kernel", s);
        VKRS_CUDA(h, launch_pdl(msd_window_kernel, dim3(1), dim3(32), 0, s, w.plan));
    }
    // the recount only works when the keys have leading zero bits (the digit window moved)
    r = msd_pass<XF>(h, w, 0, buf0, buf1, n, ctas, segments, nullptr, nullptr, 0, false, &w.plan->recount, true, false, s);
    if (r) return r;
    r = msd_pass<XF>(h, w, 0, buf0, buf1, n, ctas, segments, nullptr, w.bucket_start, 0, false, nullptr, false, true, s);
    if (r) return r;
    if (h->msd_stop_after == 1) return VKRS_OK;
    // ---- pass 2: second digit inside each of the 256 buckets ----
    {
        LaunchScope scope(h, "msd_plan_pieces_kernel", s);
        VKRS_CUDA(h, launch_pdl(msd_plan_pieces_kernel, dim3(1), dim3(MSD_PLAN_THREADS), 0, s, (const uint32_t *) w.bucket_start, (uint32_t) RADIX,
                                n, seg_keys, segments, w.pieces[1], w.seg_first[1], w.bucket_first[1], &w.plan->num_pieces[1], w.sub_start,
                                (const uint32_t *) &w.plan->skip_pass2));
    }
    r = msd_pass(h, w, 1, buf1, buf0, n, ctas, segments + RADIX, w.bucket_start, w.sub_start, (uint32_t) LOCAL_MAX, false, nullptr, true, true, s);
    if (r) return r;
    if (h->msd_stop_after == 2) return VKRS_OK;
    // ---- the (digit1, digit2) buckets, batched into items of whole buckets, sorted in shared memory, in place ----
    {
        // the item window is picked on the device (lt_window): the table is sized for the smallest one
        const uint32_t item_stride = (uint32_t) (((uint64_t) n + LT_MIN_WINDOW - 1) / LT_MIN_WINDOW) + 1;
        // (+ the list of the items msd_local_tile_kernel leaves to msd_local_redo_kernel)
        r = grow(h, h->msd_items, h->msd_items_cap, 3 * (uint64_t) item_stride, sizeof(uint32_t), false);
        if (r) return r;
        uint32_t *item_first = h->msd_items, *item_lo = h->msd_items + item_stride, *redo = h->msd_items + 2 * (size_t) item_stride;
        {
            LaunchScope scope(h, "msd_items_kernel", s);
            VKRS_CUDA(h, launch_pdl(msd_items_kernel, dim3(MSD_SUBS / 256), dim3(256), 0, s, (const uint32_t *) w.sub_start, MSD_SUBS, n,
                                    item_first, item_lo, item_stride, w.plan, w.big));
        }
        // buckets too large for shared memory ("big": msd_items_kernel listed them) are sorted by counting: a histogram of
        // their low bits, then a fill.  Both kernels exit at once when there are none.  The histogram runs BEFORE the
        // shared-memory sort: an item that fits a buffer although it holds a big bucket is sorted (and, for typed keys,
        // mapped back) by msd_local_tile_kernel, and the histogram must see the keys as pass 2 left them.
        {
            constexpr size_t big_smem = (32768 + 40) * sizeof(uint32_t);
            static thread_local int configured_device = -1;
            if (configured_device != h->device) {
                r = set_smem(h, msd_big_hist_kernel<XF>, big_smem);
                if (r) return r;
                configured_device = h->device;
            }
            LaunchScope scope(h, "msd_big_hist_kernel", s);
            VKRS_CUDA(h, launch_pdl(msd_big_hist_kernel<XF>, dim3(h->sm_count), dim3(BIG_HIST_THREADS), big_smem, s, (const uint32_t *) buf0,
                                    (const MsdPlan *) w.plan, (const BigBucket *) w.big, w.big_pool, w.big_done, w.big_vcs));
        }
        uint32_t grid = (uint32_t) (h->sm_count * VKRS_LT_MIN_BLOCKS);
        if (XF != 0) {
            static thread_local int configured_device = -1;
            if (configured_device != h->device) {
                r = set_smem(h, msd_local_tile_kernel<XF>, sizeof(LocalTileSmem));
                if (r) return r;
                configured_device = h->device;
            }
        }
        {
            LaunchScope scope(h, "msd_local_tile_kernel", s);
            VKRS_CUDA(h, launch_pdl(msd_local_tile_kernel<XF>, dim3(grid), dim3(LT_THREADS), sizeof(LocalTileSmem), s, buf0,
                                    (const uint32_t *) w.sub_start, (const uint32_t *) item_first, (const uint32_t *) item_lo, n,
                                    w.plan, h->msd_use_bins, h->debug_counters, redo));
        }
        {
            static thread_local int configured_device = -1;
            if (configured_device != h->device) {
                r = set_smem(h, msd_local_redo_kernel<XF>, sizeof(LocalRedoSmem));
                if (r) return r;
                configured_device = h->device;
            }
            LaunchScope scope(h, "msd_local_redo_kernel", s);
            VKRS_CUDA(h, launch_pdl(msd_local_redo_kernel<XF>, dim3(grid), dim3(LT_THREADS), sizeof(LocalRedoSmem), s, buf0,
                                    (const uint32_t *) w.sub_start, (const uint32_t *) item_first, (const MsdPlan *) w.plan, (const uint32_t *) redo));
        }
        // ... and the big buckets are filled from their counters last: the fill only reads the counters.
        LaunchScope scope(h, "msd_big_fill_kernel", s);
        VKRS_CUDA(h, launch_pdl(msd_big_fill_kernel<XF>, dim3(grid), dim3(512), 0, s, buf0, (const MsdPlan *) w.plan, (const BigBucket *) w.big,
                                w.big_pool, (const uint32_t *) w.big_vcs));
    }
    if (h->msd_stop_after == 3) return VKRS_OK;
    // ---- fallback: four stable LSD passes that only work if some bucket was too large ----
    for (int p = 0; p < 4; ++p) {
        uint32_t *in = (p & 1) ? buf1 : buf0, *out = (p & 1) ? buf0 : buf1;
        if (p < 3 || XF == 0)
            r = launch_seg_t<uint32_t, false, 384, 16, 2, 1>(h, in, out, nullptr, nullptr, n, 8 * p, s, 0, nullptr, true, true, nullptr,
                                                             &w.plan->fallback);
        else // typed keys: the last pass undoes the map
            r = launch_seg_t<uint32_t, false, 384, 16, 2, 1, false, false, 0, XF>(h, in, out, nullptr, nullptr, n, 8 * p, s, 0, nullptr, true, true,
                                                                                  nullptr, &w.plan->fallback);
        if (r) return r;
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

// The schedule a keys-only whole sort of n keys runs (vkrs_schedule).
int resolve_schedule(const vkrs_context *h, uint32_t n) {
    if (h->schedule != VKRS_SCHEDULE_AUTO) return h->schedule;
    if (h->variant != DEFAULT_VARIANT) return VKRS_SCHEDULE_LSD; // a tuning variant was picked: run exactly that kernel
    if (n >= AUTO_BUCKET_MIN && n <= AUTO_BUCKET_MAX) return VKRS_SCHEDULE_BUCKET;
    return n >= AUTO_UNSTABLE_FIRST_MIN ? VKRS_SCHEDULE_LSD_UNSTABLE_FIRST : VKRS_SCHEDULE_LSD;
}

int check_multi_pc(vkrs_context *h, const vkrs_multi_push_constants *pc, bool need_tiling) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!pc) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "push constants are NULL");
    if (pc->g_num_elements >= (1u << 30))
        return fail(h, VKRS_ERR_UNSUPPORTED, "g_num_elements=%u: at most 2^30-1 keys per call", pc->g_num_elements);
    if (need_tiling) {
        if (pc->g_shift > 31) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "g_shift=%u out of range", pc->g_shift);
        if (pc->g_num_elements > 0) {
            if (pc->g_num_blocks_per_workgroup == 0 || pc->g_num_workgroups == 0)
                return fail(h, VKRS_ERR_INVALID_ARGUMENT, "g_num_workgroups / g_num_blocks_per_workgroup must be > 0");
            const uint64_t covered = (uint64_t) pc->g_num_workgroups * pc->g_num_blocks_per_workgroup * 256ull;
            if (covered < pc->g_num_elements)
                return fail(h, VKRS_ERR_INVALID_ARGUMENT,
                            "tiling covers %llu keys < g_num_elements=%u (W=%u, nb=%u)", (unsigned long long) covered,
                            pc->g_num_elements, pc->g_num_workgroups, pc->g_num_blocks_per_workgroup);
        }
    }
    return VKRS_OK;
}

int staged_histograms(vkrs_context *h, const uint32_t *in, uint32_t *hist, const vkrs_multi_push_constants *pc,
                      cudaStream_t stream) {
    {
        LaunchScope scope(h, "staged_histograms_kernel", stream);
        staged_histograms_kernel<<<pc->g_num_workgroups, STAGED_HIST_THREADS, 0, stream>>>(
            in, hist, pc->g_num_elements, pc->g_shift, pc->g_num_blocks_per_workgroup);
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

int staged_scatter(vkrs_context *h, const uint32_t *in, uint32_t *out, const uint32_t *hist,
                   const vkrs_multi_push_constants *pc, const uint32_t *vin, uint32_t *vout, cudaStream_t stream) {
    const uint32_t W = pc->g_num_workgroups;
    const uint32_t chunks = (W + STAGED_CHUNK_ROWS - 1) / STAGED_CHUNK_ROWS;
    int r = grow(h, h->staged_offsets, h->staged_rows, (uint64_t) W, RADIX * sizeof(uint32_t), false);
    if (r) return r;
    r = grow(h, h->staged_chunks, h->staged_chunk_rows, (uint64_t) chunks, RADIX * sizeof(uint32_t), false);
    if (r) return r;
    {
        LaunchScope scope(h, "staged_colsum_kernel", stream);
        staged_colsum_kernel<<<chunks, 256, 0, stream>>>(hist, W, h->staged_chunks);
    }
    {
        LaunchScope scope(h, "staged_chunkscan_kernel", stream);
        staged_chunkscan_kernel<<<1, 256, 0, stream>>>(h->staged_chunks, chunks, h->staged_bin_start);
    }
    {
        LaunchScope scope(h, "staged_offsets_kernel", stream);
        staged_offsets_kernel<<<chunks, 256, 0, stream>>>(hist, W, h->staged_chunks, h->staged_bin_start,
                                                         h->staged_offsets);
    }
    VKRS_CUDA(h, cudaGetLastError());

    LaunchScope scope(h, "staged_scatter_kernel", stream);
    if (vin) {
        using Sorter = TileSorter<uint32_t, true, STAGED_THREADS, STAGED_KPT, MATCH_PTX>;
        auto kernel = staged_scatter_kernel<true, STAGED_THREADS, STAGED_KPT, MATCH_PTX, 3>;
        static thread_local int configured_device = -1;
        if (configured_device != h->device) {
            r = set_smem(h, kernel, sizeof(Sorter::Smem));
            if (r) return r;
            configured_device = h->device;
        }
        kernel<<<W, STAGED_THREADS, sizeof(Sorter::Smem), stream>>>(in, out, vin, vout, h->staged_offsets,
                                                                    pc->g_num_elements, pc->g_shift,
                                                                    pc->g_num_blocks_per_workgroup);
    } else {
        using Sorter = TileSorter<uint32_t, false, STAGED_THREADS, STAGED_KPT, MATCH_PTX>;
        auto kernel = staged_scatter_kernel<false, STAGED_THREADS, STAGED_KPT, MATCH_PTX, 4>;
        static thread_local int configured_device = -1;
        if (configured_device != h->device) {
            r = set_smem(h, kernel, sizeof(Sorter::Smem));
            if (r) return r;
            configured_device = h->device;
        }
        kernel<<<W, STAGED_THREADS, sizeof(Sorter::Smem), stream>>>(in, out, nullptr, nullptr, h->staged_offsets,
                                                                    pc->g_num_elements, pc->g_shift,
                                                                    pc->g_num_blocks_per_workgroup);
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

} // namespace

extern "C" {

const char *vkrs_version(void) { return "vkradixsort_b200 0.1 (sm_100a)"; }

uint32_t vkrs_tile_size(void) { return variant_tile(DEFAULT_VARIANT); }

uint32_t vkrs_global_invocation_size(uint32_t num_elements, uint32_t nb) {
    if (nb == 0) return 0;
    uint32_t gis = num_elements / nb; // MultiRadixSort.cpp:13-15
    if (num_elements % nb > 0) gis += 1;
    return gis;
}

uint32_t vkrs_workgroup_count(uint32_t global_invocation_size) {
    return (uint32_t) (((uint64_t) global_invocation_size + VKRS_WORKGROUP_SIZE - 1) / VKRS_WORKGROUP_SIZE); // ComputePass.h:24-29
}

int vkrs_create(vkrs_handle *out_handle, int device, uint64_t max_num_elements_hint) {
    if (!out_handle) return fail(nullptr, VKRS_ERR_INVALID_ARGUMENT, "out_handle is NULL");
    *out_handle = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, VKRS_ERR_CUDA, "no CUDA device available: %s",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count)
        return fail(nullptr, VKRS_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", device, count - 1);
    DeviceGuard guard(device);
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, VKRS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, VKRS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    vkrs_context *h = new vkrs_context();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    if (const char *v = getenv("VKRS_VARIANT")) {
        int iv = atoi(v);
        if (iv >= 0 && iv < NUM_VARIANTS) h->variant = iv;
    }
    if (const char *v = getenv("VKRS_LOCAL_BINS")) h->msd_use_bins = atoi(v) != 0 ? 1u : 0u;
    if (const char *v = getenv("VKRS_GUESS_WINDOW")) h->msd_guess_window = atoi(v) != 0;
    if (const char *v = getenv("VKRS_SCHEDULE")) {
        int iv = atoi(v);
        if (iv >= 0 && iv < VKRS_NUM_SCHEDULES) h->schedule = iv;
    }
    void *p = nullptr;
    e = cudaMalloc(&p, vkrs_context::CTRL_WORDS * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(p, 0, vkrs_context::CTRL_WORDS * sizeof(uint32_t));
    if (e == cudaSuccess) {
        h->ctrl = static_cast<uint32_t *>(p);
        e = cudaMalloc(&p, RADIX * sizeof(uint32_t));
        if (e == cudaSuccess) h->staged_bin_start = static_cast<uint32_t *>(p);
    }
    if (e != cudaSuccess) {
        int r = fail(nullptr, VKRS_ERR_CUDA, "workspace allocation failed: %s", cudaGetErrorString(e));
        vkrs_destroy(h);
        return r;
    }
    if (max_num_elements_hint > 0 && max_num_elements_hint < (1ull << 30)) {
        // Everything the default schedules need for a sort of up to `hint` keys is allocated here, so that no later
        // call has to grow a workspace (growing synchronises the device).  The tile-status arrays of the single-sweep
        // tuning variants (56 MB for 10^8 keys) are only allocated when such a variant is selected.
        const uint32_t n = (uint32_t) max_num_elements_hint;
        uint32_t ctas, segments, seg_keys;
        int r = msd_prepare(h, n, ctas, segments, seg_keys);
        if (!r) r = grow(h, h->seg_hist, h->seg_hist_rows, (uint64_t) h->sm_count * 4, RADIX * sizeof(uint32_t), false);
        if (!r) r = grow(h, h->msd_items, h->msd_items_cap, 3 * ((uint64_t) (n + LT_MIN_WINDOW - 1) / LT_MIN_WINDOW + 1), sizeof(uint32_t), false);
        if (r) {
            g_create_error = h->error;
            vkrs_destroy(h);
            return r;
        }
    }
    *out_handle = h;
    return VKRS_OK;
}

int vkrs_destroy(vkrs_handle h) {
    if (!h) return VKRS_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->ctrl);
    cudaFree(h->status[0]);
    cudaFree(h->status[1]);
    cudaFree(h->seg_hist);
    cudaFree(h->staged_offsets);
    cudaFree(h->staged_chunks);
    cudaFree(h->staged_bin_start);
    cudaFree(h->debug_counters);
    cudaFree(h->host_buf[0]);
    cudaFree(h->host_buf[1]);
    cudaFree(h->msd_ws);
    cudaFree(h->msd_items);
    for (auto e : h->host_ev)
        if (e) cudaEventDestroy(e);
    delete h;
    return VKRS_OK;
}

const char *vkrs_last_error(vkrs_handle h) { return h ? h->error.c_str() : g_create_error.c_str(); }

uint64_t vkrs_launch_count(vkrs_handle h) { return h ? h->launches : 0; }

int vkrs_get_variant(vkrs_handle h) { return h ? h->variant : -1; }

// Tuning aid: phase timers of the pipelined kernel.  enable != 0 allocates/zeroes 32 device
// counters that the kernels accumulate into; out (may be NULL) receives the current values.
int vkrs_debug_counters(vkrs_handle h, int enable, uint64_t *out) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(h->device);
    VKRS_CUDA(h, cudaDeviceSynchronize());
    if (out) {
        if (h->debug_counters) VKRS_CUDA(h, cudaMemcpy(out, h->debug_counters, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        else memset(out, 0, 32 * sizeof(uint64_t));
    }
    if (enable) {
        if (!h->debug_counters) VKRS_CUDA(h, cudaMalloc(&h->debug_counters, 32 * sizeof(uint64_t)));
        VKRS_CUDA(h, cudaMemset(h->debug_counters, 0, 32 * sizeof(uint64_t)));
    } else if (h->debug_counters) {
        cudaFree(h->debug_counters);
        h->debug_counters = nullptr;
    }
    return VKRS_OK;
}

int vkrs_set_profiling(vkrs_handle h, int enable) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(h->device);
    VKRS_CUDA(h, cudaDeviceSynchronize());
    for (auto &r : h->prof_records) {
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    h->prof_records.clear();
    h->prof_summary.clear();
    h->profiling = enable != 0;
    return VKRS_OK;
}

// Folds the event pairs recorded so far into per-kernel totals; synchronises the device.
int vkrs_profile_collect(vkrs_handle h) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(h->device);
    VKRS_CUDA(h, cudaDeviceSynchronize());
    std::map<std::string, std::pair<double, uint64_t>> acc;
    for (auto &s : h->prof_summary) acc[s.first] = s.second;
    for (auto &r : h->prof_records) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
            auto &a = acc[r.name];
            a.first += ms;
            a.second += 1;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    h->prof_records.clear();
    h->prof_summary.assign(acc.begin(), acc.end());
    return (int) h->prof_summary.size();
}

int vkrs_profile_entry(vkrs_handle h, int index, const char **name, double *total_ms, uint64_t *launches) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (index < 0 || index >= (int) h->prof_summary.size()) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "profile index %d out of range", index);
    if (name) *name = h->prof_summary[index].first.c_str();
    if (total_ms) *total_ms = h->prof_summary[index].second.first;
    if (launches) *launches = h->prof_summary[index].second.second;
    return VKRS_OK;
}

int vkrs_set_variant(vkrs_handle h, int variant) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (variant < 0 || variant >= NUM_VARIANTS) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "variant %d out of range", variant);
    h->variant = variant;
    return VKRS_OK;
}

int vkrs_set_schedule(vkrs_handle h, int schedule) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (schedule < 0 || schedule >= VKRS_NUM_SCHEDULES) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "schedule %d out of range", schedule);
    h->schedule = schedule;
    return VKRS_OK;
}

int vkrs_get_schedule(vkrs_handle h) { return h ? h->schedule : -1; }

const char *vkrs_schedule_name(int schedule) {
    switch (schedule) {
        case VKRS_SCHEDULE_AUTO: return "auto";
        case VKRS_SCHEDULE_LSD: return "lsd (4 stable passes)";
        case VKRS_SCHEDULE_LSD_UNSTABLE_FIRST: return "lsd, unstable first pass";
        case VKRS_SCHEDULE_BUCKET: return "bucket (2 unstable top-digit passes + local sort)";
        default: return "";
    }
}

// Control words of the last bucket-schedule sort: {shift1, shift2, fallback, recount, smallest key, max_sub,
// pieces of pass 1, pieces of pass 2}.  Synchronises `stream`.
int vkrs_bucket_stats(vkrs_handle h, uint32_t *out8 /* 16 words */, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!out8) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "out8 is NULL");
    static_assert(sizeof(MsdPlan) == 16 * sizeof(uint32_t), "vkrs_bucket_stats copies the plan");
    memset(out8, 0, 16 * sizeof(uint32_t));
    if (!h->msd_ws) return VKRS_OK;
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VKRS_CUDA(h, cudaMemcpyAsync(out8, h->msd_ws, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    VKRS_CUDA(h, cudaStreamSynchronize(s));
    return VKRS_OK;
}

// Hint: the keys of the following keys-only sorts lie in [lo_key, hi_key].  The bucket schedule then counts its
// first histogram in the digit window of that range (digits of key - lo_key, directly under the top bit of
// hi_key - lo_key); a wrong hint only costs the recount the schedule does anyway when the window moves.
// (0, 0xFFFFFFFF) = no hint.
int vkrs_set_key_span_hint(vkrs_handle h, uint32_t lo_key, uint32_t hi_key) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (lo_key > hi_key) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "lo_key > hi_key");
    const uint32_t x = hi_key - lo_key; // the span, as msd_window_kernel derives the digit window from (largest - smallest key)
    uint32_t top = 0;
    while (top < 31 && (x >> (top + 1)) != 0) ++top;
    h->msd_first_shift = top >= 15 ? top - 7 : 8;
    h->msd_first_base = lo_key;
    return VKRS_OK;
}

// Test aid: end the bucket schedule early so each stage can be checked on its own.
//   1 = after partition pass 1 (keys grouped by the top digit in buf1), 2 = after pass 2 (grouped by the top
//   two digits in buf0), 3 = after the local sort (no fallback passes), 0 = whole schedule.
int vkrs_debug_bucket_stop(vkrs_handle h, int stage) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (stage < 0 || stage > 3) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "stage %d out of range", stage);
    h->msd_stop_after = stage;
    return VKRS_OK;
}

int vkrs_num_variants(void) { return NUM_VARIANTS; }

const char *vkrs_variant_name(int variant) {
    return (variant >= 0 && variant < NUM_VARIANTS) ? kVariants[variant].name : "";
}

// Reads (and clears) the device-side error flag; synchronises `stream`.
int vkrs_check_device_error(vkrs_handle h, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(h->device);
    uint32_t flag = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VKRS_CUDA(h, cudaMemcpyAsync(&flag, h->ctrl + vkrs_context::CTRL_ERROR, sizeof flag, cudaMemcpyDeviceToHost, s));
    VKRS_CUDA(h, cudaStreamSynchronize(s));
    if (flag != DEVERR_NONE) {
        cudaMemsetAsync(h->ctrl + vkrs_context::CTRL_ERROR, 0, sizeof(uint32_t), s);
        // tile status may be inconsistent after an aborted look-back: start clean
        if (h->status[0]) cudaMemsetAsync(h->status[0], 0, h->status_rows * RADIX * sizeof(uint32_t), s);
        if (h->status[1]) cudaMemsetAsync(h->status[1], 0, h->status_rows * RADIX * sizeof(uint32_t), s);
        cudaStreamSynchronize(s);
        return fail(h, VKRS_ERR_INTERNAL, "device error flag %u (1 = chained-scan look-back timed out)", flag);
    }
    return VKRS_OK;
}

int vkrs_multi_histograms(vkrs_handle h, const uint32_t *elements_in, uint32_t *histograms,
                          const vkrs_multi_push_constants *pc, void *stream) {
    int r = check_multi_pc(h, pc, true);
    if (r) return r;
    if (pc->g_num_elements == 0 && pc->g_num_workgroups == 0) return VKRS_OK; // zero-size dispatch
    if (!histograms || (!elements_in && pc->g_num_elements > 0)) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    return staged_histograms(h, elements_in, histograms, pc, static_cast<cudaStream_t>(stream));
}

int vkrs_multi_scatter(vkrs_handle h, const uint32_t *elements_in, uint32_t *elements_out, const uint32_t *histograms,
                       const vkrs_multi_push_constants *pc, const uint32_t *values_in, uint32_t *values_out,
                       void *stream) {
    int r = check_multi_pc(h, pc, true);
    if (r) return r;
    if (pc->g_num_elements == 0) return VKRS_OK;
    if (!elements_in || !elements_out || !histograms) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if ((values_in == nullptr) != (values_out == nullptr))
        return fail(h, VKRS_ERR_INVALID_ARGUMENT, "values_in and values_out must both be given or both be NULL");
    DeviceGuard guard(h->device);
    return staged_scatter(h, elements_in, elements_out, histograms, pc, values_in, values_out,
                          static_cast<cudaStream_t>(stream));
}

int vkrs_multi_pass(vkrs_handle h, const uint32_t *elements_in, uint32_t *elements_out, uint32_t *histograms,
                    const vkrs_multi_push_constants *pc, void *stream) {
    int r = vkrs_multi_histograms(h, elements_in, histograms, pc, stream);
    if (r) return r;
    return vkrs_multi_scatter(h, elements_in, elements_out, histograms, pc, nullptr, nullptr, stream);
}

int vkrs_multi_sort_staged(vkrs_handle h, uint32_t *buf0, uint32_t *buf1, uint32_t *histograms,
                           const vkrs_multi_push_constants *pc_in, void *stream) {
    int r = check_multi_pc(h, pc_in, false);
    if (r) return r;
    vkrs_multi_push_constants pc = *pc_in;
    for (uint32_t i = 0; i < 4; ++i) { // MultiRadixSort.cpp:56-61
        pc.g_shift = 8 * i;
        uint32_t *in = (i & 1) ? buf1 : buf0, *out = (i & 1) ? buf0 : buf1;
        r = vkrs_multi_pass(h, in, out, histograms, &pc, stream);
        if (r) return r;
    }
    return VKRS_OK;
}

int vkrs_multi_sort(vkrs_handle h, uint32_t *buf0, uint32_t *buf1, uint32_t *histograms,
                    const vkrs_multi_push_constants *pc, void *stream) {
    (void) histograms; // the whole-sort schedules keep their histogram rows / tile status in the handle's workspace
    int r = check_multi_pc(h, pc, false);
    if (r) return r;
    const uint32_t n = pc->g_num_elements;
    if (n == 0) return VKRS_OK;
    if (!buf0 || !buf1) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sched = resolve_schedule(h, n);
    if (sched == VKRS_SCHEDULE_BUCKET) return msd_sort_t<0>(h, buf0, buf1, n, s);
    if (sched == VKRS_SCHEDULE_LSD_UNSTABLE_FIRST) return lsd_unstable_first_sort_u32(h, buf0, buf1, n, s);
    if (!variant_is_segmented(h->variant)) { // the single-sweep variants need the global digit starts
        const uint32_t tile = variant_tile(h->variant);
        r = ensure_status(h, ((uint64_t) n + tile - 1) / tile);
        if (r) return r;
        r = launch_global_histogram<uint32_t, 4>(h, buf0, n, s);
        if (r) return r;
    }
    for (int p = 0; p < 4; ++p) {
        uint32_t *in = (p & 1) ? buf1 : buf0, *out = (p & 1) ? buf0 : buf1;
        r = launch_pass_u32(h, in, out, n, 8 * p, p, s);
        if (r) return r;
    }
    return VKRS_OK;
}

int vkrs_multi_sort_typed(vkrs_handle h, void *buf0, void *buf1, uint32_t *histograms, const vkrs_multi_push_constants *pc,
                          int key_type, void *stream) {
    (void) histograms;
    int r = check_multi_pc(h, pc, false);
    if (r) return r;
    const uint32_t n = pc->g_num_elements;
    if (n == 0) return VKRS_OK;
    if (!buf0 || !buf1) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint32_t *b0 = static_cast<uint32_t *>(buf0), *b1 = static_cast<uint32_t *>(buf1);
    const bool bucket = resolve_schedule(h, n) == VKRS_SCHEDULE_BUCKET; // else: four stable LSD passes
    switch (key_type) {
        case VKRS_KEY_U32: return bucket ? msd_sort_t<0>(h, b0, b1, n, s) : typed_sort_u32<0>(h, b0, b1, n, s);
        case VKRS_KEY_I32: return bucket ? msd_sort_t<1>(h, b0, b1, n, s) : typed_sort_u32<1>(h, b0, b1, n, s);
        // float32: sign + exponent + 7 mantissa bits make the 16-bit prefix, which bell-shaped data crowds into a few
        // hundred buckets -- the fallback would be the rule; auto keeps them on the LSD passes, BUCKET can be forced
        case VKRS_KEY_F32: return h->schedule == VKRS_SCHEDULE_BUCKET ? msd_sort_t<2>(h, b0, b1, n, s) : typed_sort_u32<2>(h, b0, b1, n, s);
        default: return fail(h, VKRS_ERR_INVALID_ARGUMENT, "unknown key type %d", key_type);
    }
}

int vkrs_multi_sort_pairs(vkrs_handle h, uint32_t *keys0, uint32_t *keys1, uint32_t *values0, uint32_t *values1,
                          uint32_t *histograms, const vkrs_multi_push_constants *pc, void *stream) {
    (void) histograms;
    int r = check_multi_pc(h, pc, false);
    if (r) return r;
    const uint32_t n = pc->g_num_elements;
    if (n == 0) return VKRS_OK;
    if (!keys0 || !keys1 || !values0 || !values1) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int p = 0; p < 4; ++p) {
        uint32_t *in = (p & 1) ? keys1 : keys0, *out = (p & 1) ? keys0 : keys1;
        uint32_t *vin = (p & 1) ? values1 : values0, *vout = (p & 1) ? values0 : values1;
        r = launch_seg_t<uint32_t, true, PAIR_WORKERS, PAIR_KPT, 2, 1>(h, in, out, vin, vout, n, 8 * p, s);
        if (r) return r;
    }
    return VKRS_OK;
}

int vkrs_multi_sort_u64(vkrs_handle h, uint64_t *buf0, uint64_t *buf1, uint32_t *histograms,
                        const vkrs_multi_push_constants *pc, void *stream) {
    (void) histograms;
    int r = check_multi_pc(h, pc, false);
    if (r) return r;
    const uint32_t n = pc->g_num_elements;
    if (n == 0) return VKRS_OK;
    if (!buf0 || !buf1) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    using K = unsigned long long;
    for (int p = 0; p < 8; ++p) { // NUM_ITERATIONS = 8, MultiRadixSort.cpp:54
        K *in = reinterpret_cast<K *>((p & 1) ? buf1 : buf0), *out = reinterpret_cast<K *>((p & 1) ? buf0 : buf1);
        r = launch_seg_t<K, false, U64_WORKERS, U64_KPT, 2, 1>(h, in, out, nullptr, nullptr, n, 8 * p, s);
        if (r) return r;
    }
    return VKRS_OK;
}

int vkrs_key_range(vkrs_handle h, const uint32_t *keys, uint32_t num_elements, uint32_t *min_max_out, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!min_max_out || (!keys && num_elements > 0)) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t init[2] = {0xFFFFFFFFu, 0u};
    VKRS_CUDA(h, cudaMemcpyAsync(min_max_out, init, sizeof init, cudaMemcpyHostToDevice, s));
    if (num_elements > 0) {
        LaunchScope scope(h, "key_range_kernel", s);
        key_range_kernel<<<h->sm_count * 4, 512, 0, s>>>(keys, num_elements, min_max_out);
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

int vkrs_partition(vkrs_handle h, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *values_in,
                   uint32_t *values_out, uint32_t num_elements, uint32_t key_base, uint32_t shift,
                   uint32_t *bucket_counts, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (num_elements >= (1u << 30)) return fail(h, VKRS_ERR_UNSUPPORTED, "at most 2^30-1 keys per call");
    if (shift > 31) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "shift=%u out of range", shift);
    if (!bucket_counts) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "bucket_counts is NULL");
    if ((values_in == nullptr) != (values_out == nullptr))
        return fail(h, VKRS_ERR_INVALID_ARGUMENT, "values_in and values_out must both be given or both be NULL");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (num_elements == 0) {
        VKRS_CUDA(h, cudaMemsetAsync(bucket_counts, 0, RADIX * sizeof(uint32_t), s));
        return VKRS_OK;
    }
    if (!keys_in || !keys_out) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (values_in)
        return launch_seg_t<uint32_t, true, PAIR_WORKERS, PAIR_KPT, 2, 1, true>(h, keys_in, keys_out, values_in, values_out,
                                                                                num_elements, shift, s, key_base, bucket_counts);
    return launch_seg_t<uint32_t, false, 384, 16, 2, 1, true>(h, keys_in, keys_out, nullptr, nullptr, num_elements, shift, s,
                                                              key_base, bucket_counts);
}

int vkrs_partition_count(vkrs_handle h, const uint32_t *keys_in, uint32_t num_elements, uint32_t key_base, uint32_t shift,
                         int with_values, uint32_t *bucket_counts, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (num_elements >= (1u << 30)) return fail(h, VKRS_ERR_UNSUPPORTED, "at most 2^30-1 keys per call");
    if (shift > 31) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "shift=%u out of range", shift);
    if (!bucket_counts) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "bucket_counts is NULL");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (num_elements == 0) {
        VKRS_CUDA(h, cudaMemsetAsync(bucket_counts, 0, RADIX * sizeof(uint32_t), s));
        return VKRS_OK;
    }
    if (!keys_in) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (with_values)
        return launch_seg_t<uint32_t, true, PAIR_WORKERS, PAIR_KPT, 2, 1, true, true>(h, keys_in, nullptr, nullptr, nullptr, num_elements,
                                                                                      shift, s, key_base, bucket_counts, true, false);
    return launch_seg_t<uint32_t, false, 384, 16, 2, 1, true, true>(h, keys_in, nullptr, nullptr, nullptr, num_elements, shift, s,
                                                                    key_base, bucket_counts, true, false);
}

int vkrs_partition_scatter_p2p(vkrs_handle h, const uint32_t *keys_in, const uint32_t *values_in, uint32_t num_elements,
                               uint32_t key_base, uint32_t shift, const uint64_t *dst_tables, const uint32_t *gate, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (num_elements == 0) return VKRS_OK;
    if (num_elements >= (1u << 30)) return fail(h, VKRS_ERR_UNSUPPORTED, "at most 2^30-1 keys per call");
    if (!keys_in || !dst_tables) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned long long *tables = reinterpret_cast<const unsigned long long *>(dst_tables);
    if (values_in)
        return launch_seg_t<uint32_t, true, PAIR_WORKERS, PAIR_KPT, 2, 1, true, true>(h, keys_in, nullptr, values_in, nullptr, num_elements,
                                                                                      shift, s, key_base, nullptr, false, true, tables, gate);
    return launch_seg_t<uint32_t, false, 384, 16, 2, 1, true, true>(h, keys_in, nullptr, nullptr, nullptr, num_elements, shift, s,
                                                                    key_base, nullptr, false, true, tables, gate);
}

int vkrs_exchange_plan(vkrs_handle h, const uint32_t *all_counts, uint32_t world, uint32_t rank, const uint64_t *peer_key_ptrs,
                       const uint64_t *peer_value_ptrs, uint64_t *dst_tables, uint32_t *summary, uint32_t capacity,
                       uint32_t max_imbalance_permille, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!all_counts || !peer_key_ptrs || !dst_tables || !summary) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (world == 0 || world > (uint32_t) EXCHANGE_MAX_RANKS || rank >= world)
        return fail(h, VKRS_ERR_INVALID_ARGUMENT, "world=%u rank=%u: at most %d ranks", world, rank, EXCHANGE_MAX_RANKS);
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    LaunchScope scope(h, "exchange_plan_kernel", s);
    exchange_plan_kernel<<<1, RADIX, 0, s>>>(all_counts, world, rank, reinterpret_cast<const unsigned long long *>(peer_key_ptrs),
                                             reinterpret_cast<const unsigned long long *>(peer_value_ptrs),
                                             reinterpret_cast<unsigned long long *>(dst_tables), summary, capacity, max_imbalance_permille);
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

int vkrs_peer_barrier(vkrs_handle h, uint32_t *flags_local, const uint64_t *peer_flag_ptrs, uint32_t world, uint32_t rank, uint32_t epoch,
                      void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!flags_local || !peer_flag_ptrs) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (world == 0 || world > (uint32_t) EXCHANGE_MAX_RANKS || rank >= world)
        return fail(h, VKRS_ERR_INVALID_ARGUMENT, "world=%u rank=%u: at most %d ranks", world, rank, EXCHANGE_MAX_RANKS);
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    LaunchScope scope(h, "peer_signal_wait_kernel", s);
    peer_signal_wait_kernel<<<1, EXCHANGE_MAX_RANKS, 0, s>>>(flags_local, reinterpret_cast<const unsigned long long *>(peer_flag_ptrs), world, rank,
                                                             epoch);
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

int vkrs_ipc_alloc(vkrs_handle h, uint64_t bytes, void **device_ptr, unsigned char *ipc_handle_64) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!device_ptr || !ipc_handle_64) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL output");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    DeviceGuard guard(h->device);
    void *p = nullptr;
    VKRS_CUDA(h, cudaMalloc(&p, bytes ? bytes : 256));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(h, VKRS_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle_64, &hd, 64);
    *device_ptr = p;
    return VKRS_OK;
}

int vkrs_ipc_open(vkrs_handle h, const unsigned char *ipc_handle_64, void **device_ptr) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!device_ptr || !ipc_handle_64) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL argument");
    DeviceGuard guard(h->device);
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle_64, 64);
    VKRS_CUDA(h, cudaIpcOpenMemHandle(device_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return VKRS_OK;
}

int vkrs_ipc_close(vkrs_handle h, void *device_ptr) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(h->device);
    VKRS_CUDA(h, cudaIpcCloseMemHandle(device_ptr));
    return VKRS_OK;
}

int vkrs_ipc_free(vkrs_handle h, void *device_ptr) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(h->device);
    VKRS_CUDA(h, cudaFree(device_ptr));
    return VKRS_OK;
}

int vkrs_single_sort(vkrs_handle h, uint32_t *buf0, uint32_t *buf1, const vkrs_single_push_constants *pc,
                     void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (!pc) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "push constants are NULL");
    const uint32_t n = pc->g_num_elements;
    if (n == 0) return VKRS_OK;
    if (n >= (1u << 30)) return fail(h, VKRS_ERR_UNSUPPORTED, "g_num_elements=%u: at most 2^30-1 keys per call", n);
    if (!buf0 || !buf1) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "NULL buffer");
    DeviceGuard guard(h->device);
    using Sorter = TileSorter<uint32_t, false, SINGLE_THREADS, SINGLE_KPT, MATCH_PTX>;
    auto kernel = single_sort_kernel<SINGLE_THREADS, SINGLE_KPT, MATCH_PTX>;
    static thread_local int configured_device = -1;
    if (configured_device != h->device) {
        int r = set_smem(h, kernel, sizeof(Sorter::Smem));
        if (r) return r;
        configured_device = h->device;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n <= (uint32_t) LT_CAP - 4u) {
        // one launch, a handful of barriers: the keys are one item of the local sort (small_sort_kernel)
        static thread_local int small_configured = -1;
        if (small_configured != h->device) {
            int r = set_smem(h, small_sort_kernel, sizeof(LocalTileSmem));
            if (r) return r;
            small_configured = h->device;
        }
        LaunchScope scope(h, "small_sort_kernel", s);
        VKRS_CUDA(h, launch_pdl(small_sort_kernel, dim3(1), dim3(LT_THREADS), sizeof(LocalTileSmem), s, buf0, n));
        return VKRS_OK;
    }
    {
        LaunchScope scope(h, "single_sort_kernel", s);
        VKRS_CUDA(h, launch_pdl(kernel, dim3(1), dim3(SINGLE_THREADS), sizeof(Sorter::Smem), s, buf0, buf1, n, (const uint32_t *) nullptr));
    }
    VKRS_CUDA(h, cudaGetLastError());
    return VKRS_OK;
}

int vkrs_sort_auto(vkrs_handle h, uint32_t *buf0, uint32_t *buf1, uint32_t num_elements, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (num_elements <= AUTO_SINGLE_MAX) {
        vkrs_single_push_constants pc{num_elements};
        return vkrs_single_sort(h, buf0, buf1, &pc, stream);
    }
    vkrs_multi_push_constants pc{num_elements, 0, 0, 0};
    return vkrs_multi_sort(h, buf0, buf1, nullptr, &pc, stream);
}

int vkrs_multi_sort_host(vkrs_handle h, uint32_t *host_keys, uint32_t num_elements, void *stream) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    if (num_elements == 0) return VKRS_OK;
    if (!host_keys) return fail(h, VKRS_ERR_INVALID_ARGUMENT, "host_keys is NULL");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (num_elements > h->host_cap) {
        uint64_t c0 = h->host_cap, c1 = h->host_cap;
        int r = grow(h, h->host_buf[0], c0, num_elements, sizeof(uint32_t), false);
        if (r) return r;
        r = grow(h, h->host_buf[1], c1, num_elements, sizeof(uint32_t), false);
        if (r) return r;
        h->host_cap = num_elements;
    }
    if (!h->host_ev[0])
        for (auto &e : h->host_ev) VKRS_CUDA(h, cudaEventCreate(&e));
    const size_t bytes = (size_t) num_elements * sizeof(uint32_t);
    VKRS_CUDA(h, cudaEventRecord(h->host_ev[0], s));
    VKRS_CUDA(h, cudaMemcpyAsync(h->host_buf[0], host_keys, bytes, cudaMemcpyHostToDevice, s));
    VKRS_CUDA(h, cudaEventRecord(h->host_ev[1], s));
    int r = vkrs_sort_auto(h, h->host_buf[0], h->host_buf[1], num_elements, stream);
    if (r) return r;
    VKRS_CUDA(h, cudaEventRecord(h->host_ev[2], s));
    VKRS_CUDA(h, cudaMemcpyAsync(host_keys, h->host_buf[0], bytes, cudaMemcpyDeviceToHost, s));
    VKRS_CUDA(h, cudaEventRecord(h->host_ev[3], s));
    VKRS_CUDA(h, cudaStreamSynchronize(s));
    for (int i = 0; i < 3; ++i) {
        float ms = 0.f;
        VKRS_CUDA(h, cudaEventElapsedTime(&ms, h->host_ev[i], h->host_ev[i + 1]));
        h->host_ms[i] = ms;
    }
    return VKRS_OK;
}

int vkrs_host_timings(vkrs_handle h, double *out3) {
    if (!h || !out3) return VKRS_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 3; ++i) out3[i] = h->host_ms[i];
    return VKRS_OK;
}

int vkrs_resolve_schedule(vkrs_handle h, uint32_t num_elements) {
    if (!h) return VKRS_ERR_INVALID_ARGUMENT;
    return resolve_schedule(h, num_elements);
}

} // extern "C"
