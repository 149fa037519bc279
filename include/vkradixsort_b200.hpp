// vkradixsort_b200.hpp -- C++ facade over the C-ABI (vkradixsort_b200.h) that keeps the reference's
// class and member names, so code written against VkRadixSort's pass surface keeps compiling:
//
//   engine::GPUContext            engine/include/engine/core/GPUContext.h (only what the sort uses:
//                                 init/shutdown, getActiveIndex/incrementActiveIndex :71-73)
//   engine::Buffer                engine/include/engine/core/Buffer.h:16-113 (caller-owned device buffer,
//                                 fillDeviceWithStagingBuffer :47-62, downloadWithStagingBuffer :64-72)
//   engine::ComputePass           engine/include/engine/passes/ComputePass.h:16-60 + Pass.h:18-104
//                                 (create, release, setGlobalInvocationSize, getWorkGroupCount,
//                                 setStorageBuffer(set,binding,buf) / (frame,set,binding,buf), execute)
//   engine::MultiRadixSortPass    multiradixsort/include/MultiRadixSortPass.h:7-40
//   engine::SingleRadixSortPass   singleradixsort/include/SingleRadixSortPass.h:7-28
//
// Differences a porting maintainer has to know (all follow from Vulkan -> CUDA):
//   * no descriptor sets / pipelines / shader compilation: create() only makes a vkrs handle;
//   * a "semaphore" is stream order: execute() enqueues on the context's stream and returns a token;
//   * errors: every non-zero C-ABI status becomes std::runtime_error, as the reference throws
//     (ComputePass.h:51-53);
//   * MultiRadixSortPass::executeSort() is an addition: the whole sort in one call, free to use the
//     library's own tiling and schedule (the fast path; setSchedule(VKRS_SCHEDULE_LSD) pins the literal
//     four stable passes); execute() is the literal per-pass dispatch pair.
// Header-only; link libvkradixsort_b200.so and the CUDA runtime.
#pragma once
#include "vkradixsort_b200.h"

#include <cuda_runtime.h>

#include <cassert>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace engine {

struct Extent3D { // VkExtent3D
    uint32_t width = 0, height = 0, depth = 0;
};
using Semaphore = uint64_t; // VkSemaphore stand-in: stream order carries the dependency
constexpr Semaphore NULL_SEMAPHORE = 0;

inline void cudaCheck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

class GPUContext {
  public:
    static constexpr uint32_t MAX_FRAMES_IN_FLIGHT = 2; // GPUContext.h:110
    explicit GPUContext(int device = 0) : m_device(device) {}
    void init() {
        cudaCheck(cudaSetDevice(m_device), "cudaSetDevice");
        cudaCheck(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    }
    void shutdown() {
        if (m_stream) cudaStreamDestroy(m_stream);
        m_stream = nullptr;
    }
    [[nodiscard]] uint32_t getActiveIndex() const { return m_activeIndex; }
    void incrementActiveIndex() { m_activeIndex = (m_activeIndex + 1) % MAX_FRAMES_IN_FLIGHT; }
    [[nodiscard]] uint32_t getMultiBufferedCount() const { return MAX_FRAMES_IN_FLIGHT; }
    void waitIdle() const { cudaCheck(cudaStreamSynchronize(m_stream), "cudaStreamSynchronize"); } // vkQueueWaitIdle

    int m_device = 0;
    cudaStream_t m_stream = nullptr;

  private:
    uint32_t m_activeIndex = 0;
};

class Buffer {
  public:
    struct BufferSettings {
        uint64_t m_sizeBytes = 0; // the reference's field is uint32_t (Buffer.h:17); widened on purpose
        std::string m_name = "undefined";
    };
    Buffer(GPUContext *gpuContext, BufferSettings settings) : m_gpuContext(gpuContext), m_bufferSettings(std::move(settings)) {
        cudaCheck(cudaMalloc(&m_buffer, m_bufferSettings.m_sizeBytes ? m_bufferSettings.m_sizeBytes : 4), "cudaMalloc");
    }
    Buffer(const Buffer &) = delete;
    Buffer &operator=(const Buffer &) = delete;
    ~Buffer() { release(); }
    void release() {
        if (m_buffer) cudaFree(m_buffer);
        m_buffer = nullptr;
    }
    static std::shared_ptr<Buffer> fillDeviceWithStagingBuffer(GPUContext *gpuContext, const BufferSettings &settings, const void *data) {
        auto b = std::make_shared<Buffer>(gpuContext, settings);
        cudaCheck(cudaMemcpyAsync(b->m_buffer, data, settings.m_sizeBytes, cudaMemcpyHostToDevice, gpuContext->m_stream), "upload");
        cudaCheck(cudaStreamSynchronize(gpuContext->m_stream), "upload sync");
        return b;
    }
    void downloadWithStagingBuffer(void *data) {
        cudaCheck(cudaMemcpyAsync(data, m_buffer, m_bufferSettings.m_sizeBytes, cudaMemcpyDeviceToHost, m_gpuContext->m_stream), "download");
        cudaCheck(cudaStreamSynchronize(m_gpuContext->m_stream), "download sync");
    }
    [[nodiscard]] void *getBuffer() const { return m_buffer; }
    [[nodiscard]] uint64_t getSizeBytes() const { return m_bufferSettings.m_sizeBytes; }

  private:
    GPUContext *m_gpuContext;
    BufferSettings m_bufferSettings;
    void *m_buffer = nullptr;
};

class ComputePass {
  public:
    explicit ComputePass(GPUContext *gpuContext, uint32_t numStages) : m_gpuContext(gpuContext), m_workGroupCounts(numStages) {}
    virtual ~ComputePass() { release(); }
    // maxNumElementsHint sizes the handle's workspace up front (Pass::create builds pipelines, Pass.h:18-32)
    virtual void create(uint64_t maxNumElementsHint = 0) {
        if (vkrs_create(&m_handle, m_gpuContext->m_device, maxNumElementsHint) != VKRS_OK)
            throw std::runtime_error(std::string("Failed to create pass: ") + vkrs_last_error(nullptr));
    }
    virtual void release() {
        if (m_handle) vkrs_destroy(m_handle);
        m_handle = nullptr;
    }
    void setGlobalInvocationSize(uint32_t stageIndex, uint32_t width, uint32_t height, uint32_t depth) {
        // every shader of the reference declares local_size_x = 256, y = z = 1 (ComputePass.h:16-29)
        m_workGroupCounts.at(stageIndex) = {vkrs_workgroup_count(width), height, depth};
    }
    [[nodiscard]] Extent3D getWorkGroupCount(uint32_t stageIndex) { return m_workGroupCounts.at(stageIndex); }
    void setStorageBuffer(uint32_t set, uint32_t binding, Buffer *buffer) { // all frames, Pass.h:54-65
        for (uint32_t f = 0; f < GPUContext::MAX_FRAMES_IN_FLIGHT; f++) setStorageBuffer(f, set, binding, buffer);
    }
    void setStorageBuffer(uint32_t frame, uint32_t set, uint32_t binding, Buffer *buffer) { // Pass.h:67-104
        m_bindings[frame][{set, binding}] = buffer;
    }
    virtual Semaphore execute(Semaphore awaitBeforeExecution) = 0;
    [[nodiscard]] vkrs_handle handle() const { return m_handle; }

  protected:
    void *bound(uint32_t set, uint32_t binding) {
        auto &frame = m_bindings[m_gpuContext->getActiveIndex()];
        auto it = frame.find({set, binding});
        assert(it != frame.end() && "no storage buffer bound"); // the reference asserts too (Pass.h:59-60)
        if (it == frame.end()) throw std::runtime_error("no storage buffer bound at this set/binding");
        return it->second->getBuffer();
    }
    void check(int status) {
        if (status != VKRS_OK) throw std::runtime_error(vkrs_last_error(m_handle));
    }
    GPUContext *m_gpuContext;
    vkrs_handle m_handle = nullptr;

  private:
    std::vector<Extent3D> m_workGroupCounts;
    std::map<std::pair<uint32_t, uint32_t>, Buffer *> m_bindings[GPUContext::MAX_FRAMES_IN_FLIGHT];
};

class MultiRadixSortPass : public ComputePass {
  public:
    explicit MultiRadixSortPass(GPUContext *gpuContext) : ComputePass(gpuContext, 2) {}
    enum ComputeStage { RADIX_SORT_HISTOGRAMS = 0, RADIX_SORT = 1 }; // also the descriptor set numbers
    using PushConstantsHistograms = vkrs_multi_push_constants;       // MultiRadixSortPass.h:17-22
    using PushConstants = vkrs_multi_push_constants;                 // MultiRadixSortPass.h:26-31
    PushConstantsHistograms m_pushConstantsHistogram{};
    PushConstants m_pushConstants{};

    // One pass for the current g_shift = recordCommands (MultiRadixSortPass.cpp:10-20): the histogram
    // stage on (0,0)->(0,1), then the sort stage on (1,0),(1,2)->(1,1).
    Semaphore execute(Semaphore) override {
        cudaStream_t s = m_gpuContext->m_stream;
        check(vkrs_multi_histograms(m_handle, static_cast<const uint32_t *>(bound(RADIX_SORT_HISTOGRAMS, 0)),
                                    static_cast<uint32_t *>(bound(RADIX_SORT_HISTOGRAMS, 1)), &m_pushConstantsHistogram, s));
        check(vkrs_multi_scatter(m_handle, static_cast<const uint32_t *>(bound(RADIX_SORT, 0)),
                                 static_cast<uint32_t *>(bound(RADIX_SORT, 1)), static_cast<const uint32_t *>(bound(RADIX_SORT, 2)),
                                 &m_pushConstants, nullptr, nullptr, s));
        return ++m_token;
    }
    // Addition: which schedule executeSort() runs (vkrs_schedule; default VKRS_SCHEDULE_AUTO).
    void setSchedule(int schedule) { check(vkrs_set_schedule(m_handle, schedule)); }
    // Addition: the whole loop of MultiRadixSort::execute (MultiRadixSort.cpp:56-61) in one call on the
    // buffers bound for the active frame: (1,0) = keys in/out, (1,1) = scratch, (1,2) = histograms.
    Semaphore executeSort() {
        check(vkrs_multi_sort(m_handle, static_cast<uint32_t *>(bound(RADIX_SORT, 0)), static_cast<uint32_t *>(bound(RADIX_SORT, 1)),
                              static_cast<uint32_t *>(bound(RADIX_SORT, 2)), &m_pushConstants, m_gpuContext->m_stream));
        return ++m_token;
    }

  private:
    Semaphore m_token = 0;
};

class SingleRadixSortPass : public ComputePass {
  public:
    explicit SingleRadixSortPass(GPUContext *gpuContext) : ComputePass(gpuContext, 1) {}
    enum ComputeStage { RADIX_SORT = 0 };
    using PushConstants = vkrs_single_push_constants; // SingleRadixSortPass.h:16-18
    PushConstants m_pushConstants{};
    Semaphore execute(Semaphore) override {
        check(vkrs_single_sort(m_handle, static_cast<uint32_t *>(bound(RADIX_SORT, 0)), static_cast<uint32_t *>(bound(RADIX_SORT, 1)),
                               &m_pushConstants, m_gpuContext->m_stream));
        return ++m_token;
    }

  private:
    Semaphore m_token = 0;
};

} // namespace engine
