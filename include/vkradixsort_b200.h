/*
 * vkradixsort_b200.h -- C-ABI of the B200-native (sm_100a) 8-bit-digit radix sort that replaces
 * the VkRadixSort compute dispatches: the per-stage entry points and the key + payload / 64-bit / small sorts
 * run the reference's least-significant-digit-first loop; the keys-only whole sort may run the same building
 * blocks most-significant digit first (vkrs_set_schedule) -- every schedule returns the same bytes.
 *
 * The reference has no FFI; its seam is the C++ pass surface
 *   engine::MultiRadixSortPass  (multiradixsort/include/MultiRadixSortPass.h:7-40)
 *   engine::SingleRadixSortPass (singleradixsort/include/SingleRadixSortPass.h:7-28)
 * plus the buffer / push-constant / dispatch convention of README.md:200-241.  Each entry
 * point below names the reference interface it stands in for.  The C++ facade that keeps the
 * reference's class and member names on top of this ABI is include/vkradixsort_b200.hpp.
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  `stream` is a cudaStream_t passed as void*
 *    (NULL = the legacy default stream).
 *  - All device buffers are caller-owned (reference: Buffer objects owned by the caller,
 *    multiradixsort/src/MultiRadixSort.cpp:83-95).  Nothing is allocated per call; the
 *    handle owns a small workspace sized at vkrs_create() (grown, with a device sync, only
 *    if a later call exceeds the hint).
 *  - All work is enqueued on `stream`, no host synchronisation inside (reference: the caller
 *    blocks with vkQueueWaitIdle, MultiRadixSort.cpp:62).  The *_host entry points are the
 *    exception: they copy in, sort, copy out and return when the result is in host memory.
 *  - Result lands in buffer 0; buffer 1 and the histogram buffer hold unspecified scratch on
 *    return (README.md:132,241; MultiRadixSort.cpp:99).
 *  - Return value: VKRS_OK or a negative vkrs_status; vkrs_last_error() gives the text.  No
 *    C++ exception crosses this boundary (the facade turns failures back into
 *    std::runtime_error as the reference throws, ComputePass.h:51-53).
 *  - A handle is not thread-safe, and it serves ONE stream at a time: its workspaces (histogram rows, piece
 *    tables, plans cached per N) are only ordered by the stream of the call that wrote them, so before a handle
 *    is used on another stream the caller orders that stream behind the previous call (event / stream wait).
 *    Distinct handles on distinct streams are independent.
 */
#ifndef VKRADIXSORT_B200_H
#define VKRADIXSORT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKRS_WORKGROUP_SIZE 256u  /* multi_radixsort.comp:11 */
#define VKRS_RADIX_SORT_BINS 256u /* multi_radixsort.comp:12 */

typedef enum vkrs_status {
    VKRS_OK = 0,
    VKRS_ERR_INVALID_ARGUMENT = -1,
    VKRS_ERR_CUDA = -2,
    VKRS_ERR_UNSUPPORTED = -3, /* e.g. N >= 2^30, the reference's own uint32 byte-size cap */
    VKRS_ERR_INTERNAL = -4
} vkrs_status;

/* MultiRadixSortPass::PushConstantsHistograms / PushConstants
 * (multiradixsort/include/MultiRadixSortPass.h:17-31; GLSL side multi_radixsort.comp:17-22).
 * Same 16-byte layout; both reference structs are identical, so one type serves both. */
typedef struct vkrs_multi_push_constants {
    uint32_t g_num_elements;
    uint32_t g_shift;
    uint32_t g_num_workgroups;
    uint32_t g_num_blocks_per_workgroup;
} vkrs_multi_push_constants;

/* SingleRadixSortPass::PushConstants (singleradixsort/include/SingleRadixSortPass.h:16-18). */
typedef struct vkrs_single_push_constants {
    uint32_t g_num_elements;
} vkrs_single_push_constants;

typedef struct vkrs_context *vkrs_handle;

/* ---- lifecycle: Pass::create / Pass::release (engine/include/engine/passes/Pass.h:18-52) ----
 * No run-time shader compilation: kernels are compiled ahead of time for sm_100a. */
int vkrs_create(vkrs_handle *out_handle, int device, uint64_t max_num_elements_hint);
int vkrs_destroy(vkrs_handle handle);
const char *vkrs_last_error(vkrs_handle handle); /* handle may be NULL: last create() error */
const char *vkrs_version(void);

/* ---- dispatch sizing: ComputePass::setGlobalInvocationSize / getWorkGroupCount
 * (engine/include/engine/passes/ComputePass.h:16-29,58-60) with the global size of
 * MultiRadixSort.cpp:13-15.  Pure host arithmetic. */
uint32_t vkrs_global_invocation_size(uint32_t num_elements, uint32_t num_blocks_per_workgroup);
uint32_t vkrs_workgroup_count(uint32_t global_invocation_size); /* ceil(gis / 256) */

/* ---- per-stage entry points: the two dispatches of MultiRadixSortPass::recordCommands
 * (multiradixsort/src/MultiRadixSortPass.cpp:10-20). ------------------------------------ */

/* Stage RADIX_SORT_HISTOGRAMS = multi_radixsort_histograms.comp:31-56.
 * histograms[256*w + b] = #{keys of workgroup w's slab with digit b}; every row is fully
 * overwritten (no pre-zeroing needed).  histograms must hold g_num_workgroups*256 uint32. */
int vkrs_multi_histograms(vkrs_handle handle, const uint32_t *elements_in, uint32_t *histograms,
                          const vkrs_multi_push_constants *pc, void *stream);

/* Stage RADIX_SORT = multi_radixsort.comp:45-127: offsets from the histogram matrix, then the
 * stable rank + scatter of every workgroup slab.  values_* may be NULL (keys only, the
 * reference's behaviour); when given, a uint32 payload travels with its key (config 3). */
int vkrs_multi_scatter(vkrs_handle handle, const uint32_t *elements_in, uint32_t *elements_out,
                       const uint32_t *histograms, const vkrs_multi_push_constants *pc,
                       const uint32_t *values_in, uint32_t *values_out, void *stream);

/* One MultiRadixSortPass::execute (ComputePass.h:31-56): both stages for the current g_shift. */
int vkrs_multi_pass(vkrs_handle handle, const uint32_t *elements_in, uint32_t *elements_out,
                    uint32_t *histograms, const vkrs_multi_push_constants *pc, void *stream);

/* ---- whole sort: the timed loop of MultiRadixSort::execute (MultiRadixSort.cpp:49-62) -----
 * Sorts buf0 by 8-bit digits; result in buf0, buf1 is scratch.  pc->g_shift is ignored (the loop sets
 * 0,8,16,24, :57-58); g_num_workgroups / g_num_blocks_per_workgroup describe the caller's histogram
 * buffer (capacity g_num_workgroups*256 uint32, may be NULL) and are otherwise only validated: this
 * entry is free to use its own tiling and schedule (vkrs_set_schedule below).  Default (auto): the
 * bucket schedule for 4*10^6 .. 2.2*10^8 keys -- two partition passes on the two most significant
 * digits that rank without stability, then every 16-bit-prefix bucket sorted in shared memory --, else
 * the literal four stable passes buf0 -> buf1 -> buf0 -> buf1 -> buf0, each one segment histogram
 * kernel + one persistent, TMA-fed scatter kernel (the reference's own two-stage decomposition with
 * 2 x #SMs segments).  Same bytes in buf0 either way.  The single-sweep chained-scan ("Onesweep")
 * variants of the stable pass are selectable with vkrs_set_variant -- DESIGN.md 4.1-4.3. */
int vkrs_multi_sort(vkrs_handle handle, uint32_t *buf0, uint32_t *buf1, uint32_t *histograms,
                    const vkrs_multi_push_constants *pc, void *stream);

/* Same with a uint32 payload per key (extension, BASELINE.json config 3): stable by key. */
int vkrs_multi_sort_pairs(vkrs_handle handle, uint32_t *keys0, uint32_t *keys1, uint32_t *values0,
                          uint32_t *values1, uint32_t *histograms, const vkrs_multi_push_constants *pc,
                          void *stream);

/* The reference's compile-time SORT_64BIT variant (MultiRadixSort.h:10-18; 8 iterations,
 * MultiRadixSort.cpp:51-55). */
int vkrs_multi_sort_u64(vkrs_handle handle, uint64_t *buf0, uint64_t *buf1, uint32_t *histograms,
                        const vkrs_multi_push_constants *pc, void *stream);

/* Signed / floating-point keys.  The reference sorts unsigned keys only and tells callers to
 * preprocess negatives themselves (README.md:98-99,154-155); here the order-preserving transform is
 * fused into the first pass's reads and undone in the last pass's writes: same traffic as u32.
 * F32 order: by value, -0.0 before +0.0, NaNs at the two ends according to their sign bit. */
typedef enum vkrs_key_type { VKRS_KEY_U32 = 0, VKRS_KEY_I32 = 1, VKRS_KEY_F32 = 2 } vkrs_key_type;
int vkrs_multi_sort_typed(vkrs_handle handle, void *buf0, void *buf1, uint32_t *histograms,
                          const vkrs_multi_push_constants *pc, int key_type, void *stream);

/* The reference loop run literally through the per-stage entry points (4 x histograms +
 * scatter with the caller's tiling); kept so the staged path can be timed and compared. */
int vkrs_multi_sort_staged(vkrs_handle handle, uint32_t *buf0, uint32_t *buf1, uint32_t *histograms,
                           const vkrs_multi_push_constants *pc, void *stream);

/* ---- single-workgroup path: SingleRadixSortPass::execute with single_radixsort.comp:42-139
 * (singleradixsort/src/SingleRadixSort.cpp:12-28).  One CTA, four passes, result in buf0. */
int vkrs_single_sort(vkrs_handle handle, uint32_t *buf0, uint32_t *buf1,
                     const vkrs_single_push_constants *pc, void *stream);

/* Size-based choice between the single and the multi path (README.md:18-21 leaves this to
 * the user).  histograms may be NULL. */
int vkrs_sort_auto(vkrs_handle handle, uint32_t *buf0, uint32_t *buf1, uint32_t num_elements, void *stream);

/* ---- multi-GPU bucket exchange, device side (no reference counterpart: the reference is single
 * device; BASELINE.json config 5 defines the extension).  One process per GPU runs
 *   vkrs_key_range  -> all-reduce min/max -> pick (key_base, shift) so that
 *                      bucket(key) = min(255, (key - key_base) >> shift) spreads the occupied key
 *                      range over 256 order-preserving buckets;
 *   vkrs_partition  -> keys_out = keys_in stably grouped by bucket, bucket_counts[256] (device) = the
 *                      group sizes; contiguous bucket ranges are then dealt to the ranks, exchanged
 *                      (all-to-all-v over NVLink) and sorted locally with vkrs_multi_sort.
 * The host-side orchestration is vkradixsort_b200/dist.py. */
int vkrs_key_range(vkrs_handle handle, const uint32_t *keys, uint32_t num_elements, uint32_t *min_max_out /* device, 2 x uint32 */,
                   void *stream);
int vkrs_partition(vkrs_handle handle, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *values_in,
                   uint32_t *values_out, uint32_t num_elements, uint32_t key_base, uint32_t shift,
                   uint32_t *bucket_counts /* device, 256 x uint32 */, void *stream);

/* The same partition with the exchange FUSED into it (one kernel computes the buckets and stores
 * them straight into the owning ranks' receive buffers over NVLink peer mappings; no all-to-all):
 *   vkrs_ipc_alloc / vkrs_ipc_open     receive buffers shared between the ranks of one node (CUDA IPC)
 *   vkrs_partition_count               bucket counts only (the exchange plan is made from them)
 *   vkrs_partition_scatter_p2p         must follow vkrs_partition_count on the same input and handle.
 *                                      dst_tables (device, 1024 x uint64), per bucket b: [b] = address
 *                                      where this rank's part of the key receive buffer of b's owner
 *                                      starts, [256+b] the same for payloads (ignored without values),
 *                                      [512+b] / [768+b] = first bucket / one past the last bucket of b's
 *                                      owner (contiguous bucket ranges belong to one rank).  Inside its
 *                                      part the sender writes one run per owner and tile. */
int vkrs_ipc_alloc(vkrs_handle handle, uint64_t bytes, void **device_ptr, unsigned char *ipc_handle_64);
int vkrs_ipc_open(vkrs_handle handle, const unsigned char *ipc_handle_64, void **device_ptr);
int vkrs_ipc_close(vkrs_handle handle, void *device_ptr);
int vkrs_ipc_free(vkrs_handle handle, void *device_ptr);
int vkrs_partition_count(vkrs_handle handle, const uint32_t *keys_in, uint32_t num_elements, uint32_t key_base,
                         uint32_t shift, int with_values, uint32_t *bucket_counts /* device, 256 x uint32 */, void *stream);
int vkrs_partition_scatter_p2p(vkrs_handle handle, const uint32_t *keys_in, const uint32_t *values_in, uint32_t num_elements,
                               uint32_t key_base, uint32_t shift, const uint64_t *dst_tables,
                               const uint32_t *gate /* device word, may be NULL: the kernel only works if *gate != 0 */, void *stream);
/* The exchange's control work on the device, so that the step never returns to the host between the all-gather of the
 * bucket counts and the local sort:
 *   vkrs_exchange_plan   all_counts [world][256] (device, all-gathered) -> dst_tables for vkrs_partition_scatter_p2p
 *                        (contiguous bucket ranges dealt to the ranks as evenly as the bucket granularity allows; the
 *                        receive buffer of a rank is laid out source rank by source rank) and summary (device,
 *                        4 + world x uint32: keys this rank receives | largest receive count of any rank | gate | 0 |
 *                        boundaries[1 .. world-1]; gate = 1 iff the largest range fits `capacity` keys and, when
 *                        max_imbalance_permille != 0, largest / mean load <= permille / 1000 -- pass &summary[2] as the
 *                        gate of vkrs_partition_scatter_p2p to make the exchange conditional without a host round trip).  peer_key_ptrs / peer_value_ptrs: device arrays of the ranks'
 *                        receive-buffer addresses as this process sees them (vkrs_ipc_open); world <= 64.
 *   vkrs_peer_barrier    stream-ordered barrier between the ranks of one node through flags in peer memory:
 *                        flags_local = this rank's world x uint32 flag array (vkrs_ipc_alloc, zeroed), peer_flag_ptrs =
 *                        device array of every rank's flag array; epoch must increase by one per call on every rank. */
int vkrs_exchange_plan(vkrs_handle handle, const uint32_t *all_counts, uint32_t world, uint32_t rank, const uint64_t *peer_key_ptrs,
                       const uint64_t *peer_value_ptrs, uint64_t *dst_tables, uint32_t *summary, uint32_t capacity,
                       uint32_t max_imbalance_permille, void *stream);
int vkrs_peer_barrier(vkrs_handle handle, uint32_t *flags_local, const uint64_t *peer_flag_ptrs, uint32_t world, uint32_t rank,
                      uint32_t epoch, void *stream);

/* ---- host-buffer convenience = prepareBuffers + execute loop + verify's download
 * (MultiRadixSort.cpp:83-102): H2D of host_keys (pinned or pageable), sort on the device in
 * handle-owned buffers, D2H back into host_keys; returns after the data is back. */
int vkrs_multi_sort_host(vkrs_handle handle, uint32_t *host_keys, uint32_t num_elements, void *stream);
/* Device time (ms, CUDA events on `stream`) of the three stages of the last vkrs_multi_sort_host call on this handle:
 * out3[0] = host-to-device copy, [1] = sort, [2] = device-to-host copy -- the reference times the same region around its
 * submits only (MultiRadixSort.cpp:49-63); this is the split of the end-to-end number. */
int vkrs_host_timings(vkrs_handle handle, double *out3);

/* ---- device-side error flag (chained-scan look-back watchdog).  Synchronises `stream`,
 * returns VKRS_ERR_INTERNAL if any kernel since the last check raised the flag. ---- */
int vkrs_check_device_error(vkrs_handle handle, void *stream);

/* ---- tuning hooks: pick one of the precompiled schedules / tile configurations of the keys-only
 * whole sort (also settable with the VKRS_VARIANT environment variable at create time). */
int vkrs_resolve_schedule(vkrs_handle handle, uint32_t num_elements); /* the vkrs_schedule `auto` picks for a keys-only sort of that many keys */
int vkrs_num_variants(void);
const char *vkrs_variant_name(int variant);
int vkrs_set_variant(vkrs_handle handle, int variant);
int vkrs_get_variant(vkrs_handle handle);

/* ---- schedules of the keys-only whole sort (vkrs_multi_sort / vkrs_sort_auto / vkrs_multi_sort_host; also
 * settable with the VKRS_SCHEDULE environment variable at create time).  Every schedule produces the
 * same bytes in buf0 (a sorted uint32 array is unique, MultiRadixSort.cpp:148-161); they differ in how
 * much work per key the SMs do:
 *   LSD                  four stable 8-bit passes, least significant digit first -- the literal
 *                        MultiRadixSort::execute loop (MultiRadixSort.cpp:56-61).
 *   LSD_UNSTABLE_FIRST   the same, but pass 0 ranks keys with one shared-memory atomic instead of the
 *                        stable ballot match: a first pass has no earlier order to preserve.
 *   BUCKET               two unstable passes on the two most significant digits of (key - smallest key), i.e. of
 *                        the occupied key range, then every 16-bit-prefix bucket is sorted in shared memory.
 *                        Falls back to LSD on the device, without a host round trip, when a bucket is
 *                        larger than 4096 keys (heavily skewed input).  DESIGN.md 4.1.
 *   AUTO                 by N, from the measured crossovers: BUCKET for 4*10^6 .. 2.2*10^8 keys, LSD below,
 *                        LSD_UNSTABLE_FIRST above (LSD whenever a tuning variant other than the default was
 *                        selected with vkrs_set_variant).
 * int32 keys (vkrs_multi_sort_typed) follow the same rule; float32 keys stay on the LSD passes unless BUCKET is
 * set explicitly (bell-shaped floats crowd into few 16-bit prefixes); key+payload and 64-bit sorts always run
 * stable LSD passes. */
typedef enum vkrs_schedule {
    VKRS_SCHEDULE_AUTO = 0,
    VKRS_SCHEDULE_LSD = 1,
    VKRS_SCHEDULE_LSD_UNSTABLE_FIRST = 2,
    VKRS_SCHEDULE_BUCKET = 3,
    VKRS_NUM_SCHEDULES = 4
} vkrs_schedule;
int vkrs_set_schedule(vkrs_handle handle, int schedule);
int vkrs_get_schedule(vkrs_handle handle);
const char *vkrs_schedule_name(int schedule);
/* Hint for the BUCKET schedule: all keys of the following keys-only sorts lie in [lo_key, hi_key] (e.g. one
 * rank's key range after the multi-GPU exchange).  The schedule finds the occupied key range by itself; with
 * the hint its first histogram is already counted in the right digit window, which saves one 4 B/key recount.  A wrong
 * hint costs that recount, never correctness.  (0, 0xFFFFFFFF) = no hint: the schedule then guesses the window from
 * 16384 sample keys (right for keys that simply do not fill the 32 bits, e.g. the reference's 28-bit test keys,
 * MultiRadixSort.cpp:126; environment VKRS_GUESS_WINDOW=0 turns the guess off). */
int vkrs_set_key_span_hint(vkrs_handle handle, uint32_t lo_key, uint32_t hi_key);
/* Control words of the handle's last BUCKET sort, for tests and diagnostics: out16 = {shift of pass 1,
 * shift of pass 2 (= low bits left to the local sort), fallback taken, histogram recounted, smallest key,
 * largest bucket the shared-memory sort saw, pieces of pass 1, pieces of pass 2, largest key, 0, base of the digit
 * window, buckets finished by counting ("big" buckets: above 4096 keys), their histogram work items, 0, 0, 0}.
 * Synchronises `stream`. */
int vkrs_bucket_stats(vkrs_handle handle, uint32_t *out16, void *stream);
/* Test aid: end the BUCKET schedule after stage 1 (partition pass 1: keys grouped by the top digit, in
 * buf1), 2 (pass 2: grouped by the top two digits, in buf0) or 3 (local sort, fallback passes not
 * enqueued) so that every stage can be compared with the oracle; 0 = whole schedule (default). */
int vkrs_debug_bucket_stop(vkrs_handle handle, int stage);

/* ---- opt-in per-kernel timing (the reference only has a wall clock around the loop,
 * MultiRadixSort.cpp:49,63-65).  While enabled every kernel launch is bracketed by a CUDA event
 * pair on its stream.  vkrs_profile_collect() synchronises the device, folds the pairs into
 * per-kernel totals and returns the number of distinct kernels (or a negative status);
 * vkrs_profile_entry() reads one of them.  vkrs_set_profiling() clears the totals. */
int vkrs_set_profiling(vkrs_handle handle, int enable);
int vkrs_profile_collect(vkrs_handle handle);
int vkrs_profile_entry(vkrs_handle handle, int index, const char **name, double *total_ms, uint64_t *launches);

/* Tuning aid: per-phase cycle counters of the pipelined kernel (control warp: [0] wait for tile
 * counts, [1] claim + TMA issue, [2] look-back, [3] look-back polls, [4] tiles; worker warp 0:
 * [8..15] wait for tile, rank, barrier, digit section, wait for prefix, write-out, barrier,
 * scatter).  enable != 0 (re)starts counting, 0 stops; out (32 x uint64, may be NULL) receives the
 * totals accumulated so far.  Synchronises the device. */
int vkrs_debug_counters(vkrs_handle handle, int enable, uint64_t *out);

/* ---- introspection for tests / benches ---- */
/* Number of kernel launches the handle has enqueued since creation. */
uint64_t vkrs_launch_count(vkrs_handle handle);
/* Tile size (keys per worker group and step) of the default whole-sort schedule. */
uint32_t vkrs_tile_size(void);

#ifdef __cplusplus
}
#endif
#endif /* VKRADIXSORT_B200_H */
