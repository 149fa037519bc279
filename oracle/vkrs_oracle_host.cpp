// vkrs_oracle_host.cpp -- restatement of the reference's host-side workload generator,
// CPU baseline and verifier.  TEST INFRASTRUCTURE ONLY (see vkrs_oracle.c header).
//
//   generateRandomNumbers  multiradixsort/src/MultiRadixSort.cpp:121-133
//   sort (std::sort timer) multiradixsort/src/MultiRadixSort.cpp:141-146
//   testSort               multiradixsort/src/MultiRadixSort.cpp:148-161
// (identical twins in singleradixsort/src/SingleRadixSort.cpp:85-126)
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <parallel/algorithm>
#include <omp.h>
#endif

extern "C" {

// MultiRadixSort.cpp:121-133.  The reference seeds mt19937 from std::random_device
// (no fixed seed); any fixed seed is one draw of that.  max_value = 0x0FFFFFFF is the
// reference's 28-bit range (:126); 0xFFFFFFFF is BASELINE.json's "random uint32".
void vkrs_oracle_generate_random(uint32_t *buffer, uint64_t num_elements, uint32_t seed, uint32_t max_value) {
    std::mt19937 gen(seed);
    std::uniform_int_distribution<uint32_t> distrib(0, max_value);
    for (uint64_t i = 0; i < num_elements; i++) buffer[i] = distrib(gen);
}

void vkrs_oracle_generate_random64(uint64_t *buffer, uint64_t num_elements, uint32_t seed, uint64_t max_value) {
    std::mt19937 gen(seed);
    std::uniform_int_distribution<uint64_t> distrib(0, max_value); // :128, 0x0FFFFFFFFFFF
    for (uint64_t i = 0; i < num_elements; i++) buffer[i] = distrib(gen);
}

// MultiRadixSort.cpp:141-146: in-place single-thread std::sort, steady_clock, milliseconds.
double vkrs_oracle_std_sort(uint32_t *buffer, uint64_t num_elements) {
    auto begin = std::chrono::steady_clock::now();
    std::sort(buffer, buffer + num_elements);
    auto end = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(end - begin).count();
}

double vkrs_oracle_std_sort64(uint64_t *buffer, uint64_t num_elements) {
    auto begin = std::chrono::steady_clock::now();
    std::sort(buffer, buffer + num_elements);
    auto end = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(end - begin).count();
}

// All-cores variant of the same baseline: libstdc++
// parallel-mode sort.  Falls back to std::sort when built without -fopenmp.
double vkrs_oracle_parallel_sort(uint32_t *buffer, uint64_t num_elements) {
    auto begin = std::chrono::steady_clock::now();
#ifdef _OPENMP
    __gnu_parallel::sort(buffer, buffer + num_elements);
#else
    std::sort(buffer, buffer + num_elements);
#endif
    auto end = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(end - begin).count();
}

// The OpenMP thread count, set explicitly: launchers such as torchrun export OMP_NUM_THREADS=1, which would silently
// turn the all-host-threads reference arm of bench.py into a single-thread one.
void vkrs_oracle_set_num_threads(int threads) {
    if (threads > 0) omp_set_num_threads(threads);
}

int vkrs_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Key+payload ground truth (BASELINE.json config 3; SURVEY.md 4.3: the reference
// algorithm is a stable LSD sort, so "stable sort by key" is the determined answer).
double vkrs_oracle_stable_sort_pairs(uint32_t *keys, uint32_t *values, uint64_t num_elements) {
    auto begin = std::chrono::steady_clock::now();
    std::vector<uint64_t> idx(num_elements);
    std::iota(idx.begin(), idx.end(), uint64_t(0));
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> k(num_elements), v(num_elements);
    for (uint64_t i = 0; i < num_elements; i++) { k[i] = keys[idx[i]]; v[i] = values[idx[i]]; }
    std::copy(k.begin(), k.end(), keys);
    std::copy(v.begin(), v.end(), values);
    auto end = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(end - begin).count();
}

// MultiRadixSort.cpp:148-161.  Returns -1 when equal, -2 on size mismatch, else the
// first differing index (the reference prints it and throws "TEST FAILED.").
int64_t vkrs_oracle_test_sort(const uint32_t *reference, uint64_t reference_size, const uint32_t *out_buffer,
                              uint64_t out_size) {
    if (reference_size != out_size) return -2;
    for (uint64_t i = 0; i < reference_size; i++)
        if (reference[i] != out_buffer[i]) return (int64_t) i;
    return -1;
}

} // extern "C"
