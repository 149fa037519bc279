/*
 * vkrs_oracle.c -- CPU restatement of the VkRadixSort GLSL compute shaders.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vkradixsort_b200/ may include, link or
 * call this file.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py use it, and only as the checker / CPU arm.
 *
 * PARITY PIN: the reference ships no golden vectors, fixtures or seeded tests
 *  and cannot be built in this image (needs Vulkan + glslc at
 * run time).  The only pin the reference itself holds is the property checked by
 * MultiRadixSort::testSort (multiradixsort/src/MultiRadixSort.cpp:148-161):
 * "gpu output == std::sort(input), element-wise".  Because a sorted uint32 array is
 * unique, that property fully determines the keys-only output, so this oracle is
 * pinned by checking every function below against std::sort / numpy sort in
 * tests/test_oracle.py.  In that sense parity is pinned by the reference's own
 * acceptance test, not by golden vectors ("golden vectors: none exist").
 *
 * Every function is a *structural* restatement: the same workgroup / block / lane
 * decomposition, the same intermediate buffers (g_histograms row per workgroup,
 * ping-pong buffers), so that the CUDA stages can be compared stage by stage.
 * Workgroups are independent inside one dispatch, which is what the OpenMP
 * pragmas exploit (compiled in only with -fopenmp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WORKGROUP_SIZE 256u  /* multi_radixsort.comp:11, single_radixsort.comp:10 */
#define RADIX_SORT_BINS 256u /* multi_radixsort.comp:12 */

typedef struct {
    uint32_t g_num_elements;
    uint32_t g_shift;
    uint32_t g_num_workgroups;
    uint32_t g_num_blocks_per_workgroup;
} vkrs_oracle_push_constants; /* multiradixsort/include/MultiRadixSortPass.h:17-31 */

/* ComputePass::getDispatchSize, engine/include/engine/passes/ComputePass.h:24-29,
 * applied to the global invocation size of MultiRadixSort.cpp:13-15. */
uint32_t vkrs_oracle_global_invocation_size(uint32_t num_elements, uint32_t nb) {
    uint32_t gis = num_elements / nb;
    if (num_elements % nb > 0) gis += 1;
    return gis;
}

uint32_t vkrs_oracle_workgroup_count(uint32_t num_elements, uint32_t nb) {
    uint32_t gis = vkrs_oracle_global_invocation_size(num_elements, nb);
    return (gis + WORKGROUP_SIZE - 1) / WORKGROUP_SIZE;
}

/* multi_radixsort_histograms.comp:31-56 -- one histogram row per workgroup. */
void vkrs_oracle_multi_histograms(const uint32_t *elements_in, uint32_t *histograms,
                                  const vkrs_oracle_push_constants *pc) {
    const uint32_t n = pc->g_num_elements, shift = pc->g_shift;
    const uint32_t W = pc->g_num_workgroups, nb = pc->g_num_blocks_per_workgroup;
#pragma omp parallel for schedule(static)
    for (int64_t wID = 0; wID < (int64_t) W; wID++) {
        uint32_t histogram[RADIX_SORT_BINS]; /* shared uint[256] histogram, :29 */
        memset(histogram, 0, sizeof histogram); /* :37-40 */
        for (uint32_t index = 0; index < nb; index++) { /* :42 */
            for (uint32_t lID = 0; lID < WORKGROUP_SIZE; lID++) {
                /* :43 -- uint arithmetic in the shader; 64-bit here only to keep the
                 * guard exact for element ids that would wrap, which the reference
                 * never reaches (N < 2^30). */
                uint64_t elementId = (uint64_t) wID * nb * WORKGROUP_SIZE + (uint64_t) index * WORKGROUP_SIZE + lID;
                if (elementId < n) { /* :44 */
                    uint32_t bin = (elements_in[elementId] >> shift) & (RADIX_SORT_BINS - 1); /* :46 */
                    histogram[bin] += 1; /* :48 */
                }
            }
        }
        for (uint32_t lID = 0; lID < RADIX_SORT_BINS; lID++) /* :53-55 */
            histograms[RADIX_SORT_BINS * wID + lID] = histogram[lID];
    }
}

static uint32_t popcount32(uint32_t v) { return (uint32_t) __builtin_popcount(v); }

/* multi_radixsort.comp:45-127 -- prologue scan over all W histograms, then the
 * block-by-block bitmask ranking and scatter.  values_in/values_out are an
 * extension (BASELINE.json config 3): a payload that travels with its key; pass
 * NULL for the reference's keys-only behaviour. */
void vkrs_oracle_multi_scatter(const uint32_t *elements_in, uint32_t *elements_out,
                               const uint32_t *histograms, const vkrs_oracle_push_constants *pc,
                               const uint32_t *values_in, uint32_t *values_out) {
    const uint32_t n = pc->g_num_elements, shift = pc->g_shift;
    const uint32_t W = pc->g_num_workgroups, nb = pc->g_num_blocks_per_workgroup;

    /* Totals per bin and cross-bin exclusive scan (:56-65,74): identical for every
     * workgroup, so computed once here instead of W times. */
    uint32_t bin_total[RADIX_SORT_BINS], bin_start[RADIX_SORT_BINS];
    memset(bin_total, 0, sizeof bin_total);
    for (uint32_t j = 0; j < W; j++)
        for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) bin_total[b] += histograms[RADIX_SORT_BINS * j + b];
    uint32_t run = 0;
    for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) { bin_start[b] = run; run += bin_total[b]; }

    /* local_histogram of workgroup w = sum of rows j < w (:58-62); kept as a W x 256
     * table so the workgroups below stay independent. */
    uint32_t *local_hist = (uint32_t *) malloc((size_t) (W ? W : 1) * RADIX_SORT_BINS * sizeof(uint32_t));
    {
        uint32_t acc[RADIX_SORT_BINS];
        memset(acc, 0, sizeof acc);
        for (uint32_t j = 0; j < W; j++)
            for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) {
                local_hist[(size_t) j * RADIX_SORT_BINS + b] = acc[b];
                acc[b] += histograms[RADIX_SORT_BINS * j + b];
            }
    }

#pragma omp parallel for schedule(static)
    for (int64_t wID = 0; wID < (int64_t) W; wID++) {
        uint32_t global_offsets[RADIX_SORT_BINS]; /* :38 */
        uint32_t bin_flags[RADIX_SORT_BINS][WORKGROUP_SIZE / 32]; /* :40-43 */
        for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) /* :75-76 */
            global_offsets[b] = bin_start[b] + local_hist[(size_t) wID * RADIX_SORT_BINS + b];

        for (uint32_t index = 0; index < nb; index++) { /* :83 */
            uint64_t block_base = (uint64_t) wID * nb * WORKGROUP_SIZE + (uint64_t) index * WORKGROUP_SIZE; /* :84 */
            if (block_base >= n) break; /* every lane fails the :97 guard from here on */
            memset(bin_flags, 0, sizeof bin_flags); /* :87-91 */
            uint32_t binOffset[WORKGROUP_SIZE];
            for (uint32_t lID = 0; lID < WORKGROUP_SIZE; lID++) { /* :94-105 */
                uint64_t elementId = block_base + lID;
                if (elementId < n) {
                    uint32_t binID = (elements_in[elementId] >> shift) & (RADIX_SORT_BINS - 1);
                    binOffset[lID] = global_offsets[binID]; /* read before any :121 bump of this block */
                    bin_flags[binID][lID / 32] += 1u << (lID % 32);
                }
            }
            uint32_t bump[RADIX_SORT_BINS];
            memset(bump, 0, sizeof bump);
            for (uint32_t lID = 0; lID < WORKGROUP_SIZE; lID++) { /* :107-123 */
                uint64_t elementId = block_base + lID;
                if (elementId < n) {
                    uint32_t element_in = elements_in[elementId];
                    uint32_t binID = (element_in >> shift) & (RADIX_SORT_BINS - 1);
                    uint32_t flags_bin = lID / 32, flags_bit = 1u << (lID % 32);
                    uint32_t prefix = 0, count = 0;
                    for (uint32_t i = 0; i < WORKGROUP_SIZE / 32; i++) { /* :111-118 */
                        uint32_t bits = bin_flags[binID][i];
                        uint32_t full_count = popcount32(bits);
                        uint32_t partial_count = popcount32(bits & (flags_bit - 1));
                        prefix += (i < flags_bin) ? full_count : 0u;
                        prefix += (i == flags_bin) ? partial_count : 0u;
                        count += full_count;
                    }
                    elements_out[binOffset[lID] + prefix] = element_in; /* :119 */
                    if (values_in) values_out[binOffset[lID] + prefix] = values_in[elementId];
                    if (prefix == count - 1) bump[binID] = count; /* :120-122, applied after the block */
                }
            }
            for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) global_offsets[b] += bump[b];
        }
    }
    free(local_hist);
}

/* MultiRadixSort::execute hot section, multiradixsort/src/MultiRadixSort.cpp:12-27
 * (sizing), :34-46 (ping-pong bindings) and :56-61 (the pass loop).  Result in buf0
 * (README.md:241); buf1 and histograms hold leftovers of the last passes.
 * iterations = 4 for 32-bit keys (:51-55). */
void vkrs_oracle_multi_sort(uint32_t *buf0, uint32_t *buf1, uint32_t *histograms, uint32_t num_elements,
                            uint32_t nb, uint32_t iterations, uint32_t *val0, uint32_t *val1) {
    vkrs_oracle_push_constants pc;
    pc.g_num_elements = num_elements;
    pc.g_num_workgroups = vkrs_oracle_workgroup_count(num_elements, nb);
    pc.g_num_blocks_per_workgroup = nb;
    for (uint32_t i = 0; i < iterations; i++) {
        pc.g_shift = 8 * i; /* :57-58 */
        uint32_t *in = (i % 2 == 0) ? buf0 : buf1, *out = (i % 2 == 0) ? buf1 : buf0; /* :37-46 */
        uint32_t *vin = (i % 2 == 0) ? val0 : val1, *vout = (i % 2 == 0) ? val1 : val0;
        vkrs_oracle_multi_histograms(in, histograms, &pc);
        vkrs_oracle_multi_scatter(in, out, histograms, &pc, val0 ? vin : NULL, val0 ? vout : NULL);
    }
}

/* single_radixsort.comp:42-139 -- all four digit passes inside one workgroup,
 * ping-ponging g_elements_in / g_elements_out by iteration parity (:40,:129-133).
 * Result ends in elements_in (the last, odd iteration writes it). */
void vkrs_oracle_single_sort(uint32_t *elements_in, uint32_t *elements_out, uint32_t num_elements,
                             uint32_t *values_in, uint32_t *values_out) {
    const uint32_t n = num_elements;
    for (uint32_t iteration = 0; iteration < 4; iteration++) { /* ITERATIONS, :14,:47 */
        uint32_t shift = 8 * iteration; /* :48 */
        const uint32_t *src = (iteration % 2 == 0) ? elements_in : elements_out; /* ELEMENT_IN, :40 */
        uint32_t *dst = (iteration % 2 == 0) ? elements_out : elements_in;       /* :129-133 */
        const uint32_t *vsrc = (iteration % 2 == 0) ? values_in : values_out;
        uint32_t *vdst = (iteration % 2 == 0) ? values_out : values_in;

        uint32_t histogram[RADIX_SORT_BINS]; /* :30 */
        memset(histogram, 0, sizeof histogram); /* :50-54 */
        for (uint32_t ID = 0; ID < n; ID++) /* :56-61, every lane's strided loop */
            histogram[(src[ID] >> shift) & (RADIX_SORT_BINS - 1)] += 1;

        uint32_t global_offsets[RADIX_SORT_BINS]; /* :65-84: exclusive scan over bins */
        uint32_t offset = 0;
        for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) { global_offsets[b] = offset; offset += histogram[b]; }

        uint32_t bin_flags[RADIX_SORT_BINS][WORKGROUP_SIZE / 32];
        for (uint32_t blockID = 0; blockID < n; blockID += WORKGROUP_SIZE) { /* :91 */
            memset(bin_flags, 0, sizeof bin_flags); /* :97-101 */
            uint32_t binOffset[WORKGROUP_SIZE];
            for (uint32_t lID = 0; lID < WORKGROUP_SIZE; lID++) { /* :104-115 */
                uint32_t ID = blockID + lID;
                if (ID < n) {
                    uint32_t binID = (src[ID] >> shift) & (RADIX_SORT_BINS - 1);
                    binOffset[lID] = global_offsets[binID];
                    bin_flags[binID][lID / 32] += 1u << (lID % 32);
                }
            }
            uint32_t bump[RADIX_SORT_BINS];
            memset(bump, 0, sizeof bump);
            for (uint32_t lID = 0; lID < WORKGROUP_SIZE; lID++) { /* :117-137 */
                uint32_t ID = blockID + lID;
                if (ID < n) {
                    uint32_t element_in = src[ID];
                    uint32_t binID = (element_in >> shift) & (RADIX_SORT_BINS - 1);
                    uint32_t flags_bin = lID / 32, flags_bit = 1u << (lID % 32);
                    uint32_t prefix = 0, count = 0;
                    for (uint32_t i = 0; i < WORKGROUP_SIZE / 32; i++) {
                        uint32_t bits = bin_flags[binID][i];
                        uint32_t full_count = popcount32(bits);
                        uint32_t partial_count = popcount32(bits & (flags_bit - 1));
                        prefix += (i < flags_bin) ? full_count : 0u;
                        prefix += (i == flags_bin) ? partial_count : 0u;
                        count += full_count;
                    }
                    dst[binOffset[lID] + prefix] = element_in;
                    if (values_in) vdst[binOffset[lID] + prefix] = vsrc[ID];
                    if (prefix == count - 1) bump[binID] = count;
                }
            }
            for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) global_offsets[b] += bump[b];
        }
    }
}

/* 64-bit key variant of the multi path: the reference's compile-time SORT_64BIT
 * switch (multiradixsort/include/MultiRadixSort.h:10-18) = the same two shaders
 * with uint64_t element buffers and 8 iterations (MultiRadixSort.cpp:51-55).
 * Restated compactly: per pass, per-workgroup histogram rows + stable scatter. */
void vkrs_oracle_multi_sort64(uint64_t *buf0, uint64_t *buf1, uint32_t *histograms, uint32_t num_elements,
                              uint32_t nb, uint32_t iterations) {
    const uint32_t n = num_elements;
    const uint32_t W = vkrs_oracle_workgroup_count(n, nb);
    const uint64_t slab = (uint64_t) nb * WORKGROUP_SIZE;
    for (uint32_t it = 0; it < iterations; it++) {
        uint32_t shift = 8 * it;
        uint64_t *in = (it % 2 == 0) ? buf0 : buf1, *out = (it % 2 == 0) ? buf1 : buf0;
        for (uint32_t w = 0; w < W; w++) {
            uint32_t *row = histograms + (size_t) w * RADIX_SORT_BINS;
            memset(row, 0, RADIX_SORT_BINS * sizeof(uint32_t));
            uint64_t lo = (uint64_t) w * slab, hi = lo + slab < n ? lo + slab : n;
            for (uint64_t e = lo; e < hi; e++) row[(in[e] >> shift) & 255u] += 1;
        }
        uint32_t total[RADIX_SORT_BINS], start[RADIX_SORT_BINS];
        memset(total, 0, sizeof total);
        for (uint32_t w = 0; w < W; w++)
            for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) total[b] += histograms[(size_t) w * RADIX_SORT_BINS + b];
        uint32_t run = 0;
        for (uint32_t b = 0; b < RADIX_SORT_BINS; b++) { start[b] = run; run += total[b]; }
        /* workgroups in order, lanes in order == the shader's offsets (see multi_scatter) */
        for (uint64_t e = 0; e < n; e++) out[start[(in[e] >> shift) & 255u]++] = in[e];
    }
}
