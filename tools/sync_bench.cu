// Micro-benchmark of the synchronisation primitives the tcgen05 conv kernel chains per tile:
//  (1) mbarrier ping-pong between two warps (arrive -> try_wait wake-up), with / without suspend hint
//  (2) tcgen05.commit -> mbarrier completion latency, with 0 / 6 MMAs in front of it
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int HINT>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        if (HINT)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

template <int HINT>
__global__ void pingpong(int iters, long long* out) {
    __shared__ uint64_t bars[2];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
    const long long t0 = clock64();
    if (lane == 0) {
        for (int i = 0; i < iters; ++i) {
            if (warp == 0) { arrive(b0); wait<HINT>(b1, i & 1); }
            else if (warp == 1) { wait<HINT>(b0, i & 1); arrive(b1); }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

__global__ void __launch_bounds__(128, 1) commit_lat(int nmma, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint64_t ad = ((uint64_t)hi << 32) | ((smem_u32(smem) >> 4) | ((2048u >> 4) << 16));
        const uint64_t bd = ((uint64_t)hi << 32) | (((smem_u32(smem) + 32768) >> 4) | ((512u >> 4) << 16));
        long long tot = 0;
        for (int it = 0; it < iters; ++it) {
            const long long t0 = clock64();
            for (int q = 0; q < nmma; ++q)
                asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            wait<0>(smem_u32(&bar), it & 1);
            tot += clock64() - t0;
        }
        if (blockIdx.x == 0) out[0] = tot;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        pingpong<0><<<148, 64>>>(iters, out); cudaDeviceSynchronize();
        if (rep) printf("mbarrier ping-pong, plain try_wait : %.1f cycles per round trip (2 hops)\n", (double)out[0] / iters);
        pingpong<1><<<148, 64>>>(iters, out); cudaDeviceSynchronize();
        if (rep) printf("mbarrier ping-pong, suspend hint   : %.1f cycles per round trip (2 hops)\n", (double)out[0] / iters);
    }
    cudaFuncSetAttribute(commit_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (int nmma : {0, 1, 6, 18}) {
        commit_lat<<<148, 128, 128 * 1024>>>(nmma, 500, out); cudaDeviceSynchronize();
        commit_lat<<<148, 128, 128 * 1024>>>(nmma, 500, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        printf("tcgen05: %2d MMAs (N=32) + commit + wait : %.1f cycles\n", nmma, (double)out[0] / 500);
    }
    return 0;
}
