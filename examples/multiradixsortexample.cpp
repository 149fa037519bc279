// multiradixsortexample -- the reference's multi-radix-sort example program on the B200 library.
//
// Same program logic, stdout lines and exit codes as multiradixsort/src/bin/MultiRadixSortExample.cpp +
// MultiRadixSort::execute (multiradixsort/src/MultiRadixSort.cpp:5-81): generate keys, upload, size the
// dispatch, bind the ping-pong buffers, run the four digit passes, time them, std::sort on the CPU,
// download buffer 0 and compare element by element.
//
//   multiradixsortexample [N=1000000] [nb=32] [--fast] [--seed S] [--bits 28|32] [--csv FILE] [--reps R]
//     --fast   run the whole-sort entry (MultiRadixSortPass::executeSort, the library's own tiling)
//              instead of the literal per-pass dispatch pairs
//     --seed   fixed mt19937 seed (default: std::random_device, as the reference, :123-124)
//     --bits   key range: 28 = the reference's uniform(0, 0x0FFFFFFF) (:126), 32 = full range
//     --csv    append "N nb gpu_ms cpu_ms" (the line the reference keeps commented out, :78-80)
#include "../include/vkradixsort_b200.hpp"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>

namespace {

const char *PRINT_PREFIX = "[MultiRadixSort] ";

struct Options {
    uint32_t numElements = 1000000; // MultiRadixSort.h:29
    uint32_t numBlocksPerWorkgroup = 32; // MultiRadixSort.cpp:12
    bool fast = false, haveSeed = false;
    uint32_t seed = 0, bits = 28, reps = 1;
    std::string csv;
};

double millisSince(std::chrono::steady_clock::time_point begin) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - begin).count();
}

void run(engine::GPUContext *gpu, const Options &opt) {
    using engine::MultiRadixSortPass;
    const uint32_t N = opt.numElements, nb = opt.numBlocksPerWorkgroup;

    // ---- workload (generateRandomNumbers, :121-133) ----
    std::vector<uint32_t> elementsIn(N);
    {
        std::random_device rd;
        std::mt19937 gen(opt.haveSeed ? opt.seed : rd());
        std::uniform_int_distribution<uint32_t> distrib(0, opt.bits >= 32 ? 0xFFFFFFFFu : 0x0FFFFFFFu);
        for (auto &e : elementsIn) e = distrib(gen);
    }

    // ---- pass + dispatch sizing (:10-27) ----
    auto pass = std::make_shared<MultiRadixSortPass>(gpu);
    pass->create(N);
    uint32_t globalInvocationSize = N / nb + (N % nb > 0 ? 1 : 0);
    pass->setGlobalInvocationSize(MultiRadixSortPass::RADIX_SORT_HISTOGRAMS, globalInvocationSize, 1, 1);
    pass->setGlobalInvocationSize(MultiRadixSortPass::RADIX_SORT, globalInvocationSize, 1, 1);
    const uint32_t numWorkgroups = pass->getWorkGroupCount(MultiRadixSortPass::RADIX_SORT_HISTOGRAMS).width;
    for (auto *pc : {&pass->m_pushConstantsHistogram, &pass->m_pushConstants}) {
        pc->g_num_elements = N;
        pc->g_num_workgroups = numWorkgroups;
        pc->g_num_blocks_per_workgroup = nb;
    }

    // ---- buffers (prepareBuffers, :83-95): keys, scratch, histograms ----
    const uint64_t bytes = uint64_t(N) * sizeof(uint32_t);
    std::vector<uint32_t> zeros(std::max<uint64_t>(N, uint64_t(numWorkgroups) * 256), 0u);
    std::vector<std::shared_ptr<engine::Buffer>> buffers(3);
    buffers[1] = engine::Buffer::fillDeviceWithStagingBuffer(gpu, {bytes, "radixsort.elements1"}, zeros.data());
    buffers[2] = engine::Buffer::fillDeviceWithStagingBuffer(gpu, {uint64_t(std::max(1u, numWorkgroups)) * 256 * sizeof(uint32_t), "radixsort.histograms"}, zeros.data());
    std::cout << PRINT_PREFIX << "Sorting " << N << " " << (sizeof(elementsIn[0]) * 8) << "bit numbers." << std::endl;

    double gpuSortTime = 0;
    for (uint32_t rep = 0; rep < opt.reps; rep++) {
        buffers[0] = engine::Buffer::fillDeviceWithStagingBuffer(gpu, {bytes, "radixsort.elements0"}, elementsIn.data());
        // ---- ping-pong bindings (:34-46) ----
        const uint32_t a = gpu->getActiveIndex(), b = (a + 1) % 2;
        pass->setStorageBuffer(a, MultiRadixSortPass::RADIX_SORT_HISTOGRAMS, 0, buffers[0].get());
        pass->setStorageBuffer(a, MultiRadixSortPass::RADIX_SORT, 0, buffers[0].get());
        pass->setStorageBuffer(b, MultiRadixSortPass::RADIX_SORT, 1, buffers[0].get());
        pass->setStorageBuffer(b, MultiRadixSortPass::RADIX_SORT_HISTOGRAMS, 0, buffers[1].get());
        pass->setStorageBuffer(a, MultiRadixSortPass::RADIX_SORT, 1, buffers[1].get());
        pass->setStorageBuffer(b, MultiRadixSortPass::RADIX_SORT, 0, buffers[1].get());
        pass->setStorageBuffer(MultiRadixSortPass::RADIX_SORT_HISTOGRAMS, 1, buffers[2].get());
        pass->setStorageBuffer(MultiRadixSortPass::RADIX_SORT, 2, buffers[2].get());

        // ---- the timed region (:49-63) ----
        auto begin = std::chrono::steady_clock::now();
        if (opt.fast) {
            pass->executeSort();
        } else {
            engine::Semaphore await = engine::NULL_SEMAPHORE;
            for (uint32_t i = 0; i < 4; i++) { // NUM_ITERATIONS, SORT_32BIT
                pass->m_pushConstantsHistogram.g_shift = 8 * i;
                pass->m_pushConstants.g_shift = 8 * i;
                await = pass->execute(await);
                gpu->incrementActiveIndex();
            }
        }
        gpu->waitIdle();
        gpuSortTime = millisSince(begin);
        std::cout << PRINT_PREFIX << "GPU sort finished in " << gpuSortTime << "[ms]." << std::endl;
    }

    // ---- CPU baseline (sort, :141-146) ----
    auto begin = std::chrono::steady_clock::now();
    std::sort(elementsIn.begin(), elementsIn.end());
    const double cpuSortTime = millisSince(begin);
    std::cout << PRINT_PREFIX << "CPU sort finished in " << cpuSortTime << "[ms]." << std::endl;

    // ---- verify (:97-102, testSort :148-161): the result is in buffer 0 ----
    std::vector<uint32_t> out(N);
    buffers[0]->downloadWithStagingBuffer(out.data());
    if (out.size() != elementsIn.size()) {
        std::cerr << PRINT_PREFIX << "reference.size() != outBuffer.size()" << std::endl;
        throw std::runtime_error("TEST FAILED.");
    }
    for (uint32_t i = 0; i < N; i++) {
        if (elementsIn[i] != out[i]) {
            std::cerr << PRINT_PREFIX << elementsIn[i] << " = reference[" << i << "] != outBuffer[" << i << "] = " << out[i] << std::endl;
            throw std::runtime_error("TEST FAILED.");
        }
    }
    std::cout << PRINT_PREFIX << "Test passed." << std::endl;

    if (!opt.csv.empty()) {
        std::ofstream f(opt.csv, std::ios_base::app);
        f << N << " " << nb << " " << gpuSortTime << " " << cpuSortTime << std::endl;
    }
    for (auto &b : buffers) b->release();
    pass->release();
}

} // namespace

int main(int argc, char **argv) {
    Options opt;
    int positional = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--fast") opt.fast = true;
        else if (a == "--seed") { opt.seed = (uint32_t) std::strtoul(next(), nullptr, 0); opt.haveSeed = true; }
        else if (a == "--bits") opt.bits = (uint32_t) std::atoi(next());
        else if (a == "--csv") opt.csv = next();
        else if (a == "--reps") opt.reps = std::max(1, std::atoi(next()));
        else if (positional == 0) { opt.numElements = (uint32_t) std::strtod(a.c_str(), nullptr); positional++; }
        else if (positional == 1) { opt.numBlocksPerWorkgroup = std::max(1, std::atoi(a.c_str())); positional++; }
    }
    engine::GPUContext gpu(0);
    try {
        gpu.init();
        run(&gpu, opt);
        gpu.shutdown();
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
