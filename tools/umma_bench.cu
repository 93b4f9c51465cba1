// Micro-benchmark: issue rate / latency of tcgen05.mma (kind::f16, M=128, K=16, SS) as a function of N
// and of how many independent TMEM accumulators the stream of MMAs rotates over.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int count, int same_smem, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // finite bf16 values
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint32_t a0 = (smem_u32(smem) >> 4) | ((2048u >> 4) << 16);
        const uint32_t b0 = ((smem_u32(smem) + 32768) >> 4) | (((uint32_t)N * 16u >> 4) << 16);
        uint64_t ad[8];
        uint32_t dd[8];
        for (int q = 0; q < 8; ++q) {
            ad[q] = ((uint64_t)hi << 32) | (a0 + (same_smem ? 0u : (uint32_t)(q * 256)));
            dd[q] = tmem + (uint32_t)((q % nacc) * N);
        }
        const uint64_t bd = ((uint64_t)hi << 32) | b0;
        // first pass initialises the accumulators (no accumulate), timed passes accumulate
#pragma unroll
        for (int q = 0; q < 8; ++q)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(dd[q]), "l"(ad[q]), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        const long long t0 = clock64();
        for (int i = 0; i < count; i += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;"
                             ::"r"(dd[q]), "l"(ad[q]), "l"(bd), "r"(idesc) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            if (clock64() - t1 > 2000000000LL) break;
        }
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    const int count = 512;
    printf("%6s %5s %5s %12s %12s\n", "N", "nacc", "same", "issue cyc/mma", "total cyc/mma");
    for (int same = 0; same < 2; ++same)
        for (int N : {16, 32, 64, 96, 128, 256})
            for (int nacc : {1, 2, 4}) {
                if (nacc * N > 512) continue;
                k<<<148, 128, 128 * 1024>>>(N, nacc, count, same, out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                k<<<148, 128, 128 * 1024>>>(N, nacc, count, same, out);
                cudaDeviceSynchronize();
                printf("%6d %5d %5d %12.1f %12.1f\n", N, nacc, same, (double)out[0] / count, (double)out[1] / count);
            }
    return 0;
}
