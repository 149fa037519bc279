// Micro-benchmark: how many small 1-D TMA bulk stores (shared -> global) per second can the threads of a CTA issue?
// (Question behind it: can the scatter's write-out -- 256 runs of ~30 keys per 7680-key tile -- leave as one bulk copy
// per run instead of one STG per key?)   nvcc -O3 -gencode arch=compute_100a,code=sm_100a bulk_store_rate.cu -o bulk_store_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t bytes, uint32_t rounds, uint32_t issuers) {
    extern __shared__ __align__(128) uint32_t sm[];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < 256 * 64; i += 256) sm[i] = i + blockIdx.x;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // thread t owns "run" t: 256 B of shared memory per thread, destination stream of its own
    const size_t stream = ((size_t) blockIdx.x * 256 + tid) * (size_t) (rounds * 256 / 4);
    if (MODE == 0) { // bulk stores
        if (tid < issuers) {
            for (uint32_t r = 0; r < rounds; ++r) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + stream + r * 64), "r"(smem_u32(sm + tid * 64)),
                             "r"(bytes)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if ((r & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else { // the same bytes with one 4-byte store per lane: warp w writes run after run
        const uint32_t lane = tid & 31, warp = tid >> 5;
        const uint32_t words = bytes / 4;
        for (uint32_t r = 0; r < rounds; ++r)
            for (uint32_t run = warp; run < issuers; run += 8) {
                const size_t st = ((size_t) blockIdx.x * 256 + run) * (size_t) (rounds * 256 / 4);
                if (lane < words) out[st + r * 64 + lane] = sm[run * 64 + lane];
            }
    }
}

int main() {
    uint32_t *out;
    const uint32_t rounds = 512;
    const size_t words = (size_t) 148 * 256 * (rounds * 256 / 4);
    cudaMalloc(&out, words * 4);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (uint32_t issuers : {32u, 256u})
            for (uint32_t bytes : {16u, 48u, 96u, 128u, 256u}) {
                float best = 1e9f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0) k<0><<<148, 256, 65536>>>(out, bytes, rounds, issuers);
                    else k<1><<<148, 256, 65536>>>(out, bytes, rounds, issuers);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    best = ms < best ? ms : best;
                }
                const double ops = 148.0 * issuers * rounds;
                printf("{\"mode\": \"%s\", \"issuers\": %u, \"bytes\": %u, \"ms\": %.4f, \"ops_per_us_per_sm\": %.1f, \"GBps\": %.0f, \"err\": \"%s\"}\n",
                       mode == 0 ? "bulk" : "stg", issuers, bytes, best, ops / 148.0 / (best * 1e3), ops * bytes / (best * 1e6),
                       cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
