// Where does the fixed cost of one small kernel in a CUDA-graph chain go?  Chains of N dependent launches, replayed as
// a graph, time per launch:
//   A  empty kernel, 256 threads, no smem                           (pure dependent-launch floor)
//   B  empty kernel, 544 threads, 165 KB dynamic smem               (CTA footprint of tc_conv)
//   C  B + mbarrier init + __syncthreads                            (barrier set-up)
//   D  C + tcgen05.alloc / dealloc of 512 TMEM columns              (TMEM set-up)
//   E  alternating A and D                                          (smem carve-out switching between kernel types)
//   F  D with 64 TMEM columns
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_bench tools/launch_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_empty(int* sink) { if (sink && threadIdx.x == 9999) *sink = 1; }
__global__ void __launch_bounds__(544, 1) k_big(int* sink) {
    extern __shared__ uint8_t sm[];
    if (sink && threadIdx.x == 9999) *sink = sm[0];
}
template <int COLS, bool TMEM>
__global__ void __launch_bounds__(544, 1) k_setup(int* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int warp = threadIdx.x >> 5;
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 1024);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 20; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(sm + 8 * i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (TMEM && warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = *slot;
    if (sink && threadIdx.x == 9999) *sink = t;
    __syncthreads();
    if (TMEM && warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(COLS) : "memory");
}
template <typename F>
float chain(F launch, int n, cudaStream_t s) {
    cudaGraph_t g; cudaGraphExec_t e;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < n; ++i) launch(i);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&e, g, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) cudaGraphLaunch(e, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int i = 0; i < 10; ++i) cudaGraphLaunch(e, s);
    cudaEventRecord(b, s);
    cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaGraphExecDestroy(e); cudaGraphDestroy(g);
    return ms * 1000.f / (10.f * n);
}
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    const int big = 165 * 1024;
    cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_setup<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_setup<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_setup<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    const int n = 64;
    for (int grid : {36, 148}) {
        printf("grid %d\n", grid);
        printf("  A empty 256thr            : %.2f us/launch\n", chain([&](int) { k_empty<<<grid, 256, 0, s>>>(nullptr); }, n, s));
        printf("  B 544thr 165KB smem       : %.2f us/launch\n", chain([&](int) { k_big<<<grid, 544, big, s>>>(nullptr); }, n, s));
        printf("  B' 544thr 48KB smem       : %.2f us/launch\n", chain([&](int) { k_big<<<grid, 544, 48 * 1024, s>>>(nullptr); }, n, s));
        printf("  C + mbarrier init, sync   : %.2f us/launch\n", chain([&](int) { k_setup<512, false><<<grid, 544, big, s>>>(nullptr); }, n, s));
        printf("  D + TMEM alloc 512 cols   : %.2f us/launch\n", chain([&](int) { k_setup<512, true><<<grid, 544, big, s>>>(nullptr); }, n, s));
        printf("  F + TMEM alloc 64 cols    : %.2f us/launch\n", chain([&](int) { k_setup<64, true><<<grid, 544, big, s>>>(nullptr); }, n, s));
        printf("  E alternating A / D       : %.2f us/launch\n", chain([&](int i) { if (i & 1) k_setup<512, true><<<grid, 544, big, s>>>(nullptr); else k_empty<<<grid * 8, 256, 0, s>>>(nullptr); }, n, s));
    }
    printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
