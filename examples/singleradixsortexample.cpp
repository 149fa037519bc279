// singleradixsortexample -- the reference's single-work-group example on the B200 library
// (singleradixsort/src/bin/SingleRadixSortExample.cpp + SingleRadixSort::execute,
// singleradixsort/src/SingleRadixSort.cpp:5-47): one dispatch, four passes inside the kernel,
// result in buffer 0.
//
//   singleradixsortexample [N=1000000] [--seed S] [--bits 28|32] [--csv FILE]
#include "../include/vkradixsort_b200.hpp"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <random>

namespace {
const char *PRINT_PREFIX = "[SingleRadixSort] ";
}

int main(int argc, char **argv) {
    uint32_t N = 1000000, seed = 0, bits = 28; // SingleRadixSort.h:30
    bool haveSeed = false;
    std::string csv;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--seed") { seed = (uint32_t) std::strtoul(next(), nullptr, 0); haveSeed = true; }
        else if (a == "--bits") bits = (uint32_t) std::atoi(next());
        else if (a == "--csv") csv = next();
        else N = (uint32_t) std::strtod(a.c_str(), nullptr);
    }
    engine::GPUContext gpu(0);
    try {
        gpu.init();
        std::vector<uint32_t> elementsIn(N);
        {
            std::random_device rd;
            std::mt19937 gen(haveSeed ? seed : rd());
            std::uniform_int_distribution<uint32_t> distrib(0, bits >= 32 ? 0xFFFFFFFFu : 0x0FFFFFFFu);
            for (auto &e : elementsIn) e = distrib(gen);
        }
        using engine::SingleRadixSortPass;
        auto pass = std::make_shared<SingleRadixSortPass>(&gpu);
        pass->create(N);
        pass->setGlobalInvocationSize(SingleRadixSortPass::RADIX_SORT, 256, 1, 1); // one work group, :12
        pass->m_pushConstants.g_num_elements = N;                                   // :15
        const uint64_t bytes = uint64_t(N) * sizeof(uint32_t);
        std::vector<uint32_t> zeros(N, 0u);
        auto buf0 = engine::Buffer::fillDeviceWithStagingBuffer(&gpu, {bytes, "radixsort.elements0"}, elementsIn.data());
        auto buf1 = engine::Buffer::fillDeviceWithStagingBuffer(&gpu, {bytes, "radixsort.elements1"}, zeros.data());
        std::cout << PRINT_PREFIX << "Sorting " << N << " " << (sizeof(elementsIn[0]) * 8) << "bit numbers." << std::endl;
        pass->setStorageBuffer(SingleRadixSortPass::RADIX_SORT, 0, buf0.get()); // :22-23
        pass->setStorageBuffer(SingleRadixSortPass::RADIX_SORT, 1, buf1.get());

        auto begin = std::chrono::steady_clock::now();
        pass->execute(engine::NULL_SEMAPHORE);
        gpu.waitIdle();
        const double gpuSortTime = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - begin).count();
        std::cout << PRINT_PREFIX << "GPU sort finished in " << gpuSortTime << "[ms]." << std::endl;

        begin = std::chrono::steady_clock::now();
        std::sort(elementsIn.begin(), elementsIn.end());
        const double cpuSortTime = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - begin).count();
        std::cout << PRINT_PREFIX << "CPU sort finished in " << cpuSortTime << "[ms]." << std::endl;

        std::vector<uint32_t> out(N);
        buf0->downloadWithStagingBuffer(out.data());
        for (uint32_t i = 0; i < N; i++) {
            if (elementsIn[i] != out[i]) {
                std::cerr << PRINT_PREFIX << elementsIn[i] << " = reference[" << i << "] != outBuffer[" << i << "] = " << out[i] << std::endl;
                throw std::runtime_error("TEST FAILED.");
            }
        }
        std::cout << PRINT_PREFIX << "Test passed." << std::endl;
        if (!csv.empty()) {
            std::ofstream f(csv, std::ios_base::app);
            f << N << " " << gpuSortTime << " " << cpuSortTime << std::endl;
        }
        buf0->release();
        buf1->release();
        pass->release();
        gpu.shutdown();
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
