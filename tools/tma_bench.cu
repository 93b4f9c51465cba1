// How fast can one SM pull a conv window of the chunk-major operand into shared memory?  Per CTA: a ring of 4 slots,
// one elected thread issues the loads of a window (3 chunks x 130 rows x 16 B = 6.2 KB, the C = 24 full-rate case, or
// 8 chunks x 128 rows = 16 KB), another waits on the slot's mbarrier; windows advance through a 64 MB plane.
//   A  cp.async.bulk.tensor.3d, view {8 elem, rows, chunks}, box {8, R, C}        (16-byte inner rows, any start row)
//   B  cp.async.bulk.tensor.3d, view {64 elem, rows/8, chunks}, box {64, R/8, C}   (128-byte inner rows, start % 8 == 0)
//   C  1-D cp.async.bulk per chunk (R*16 bytes each)
//   D  per-thread cp.async 16 B (128 threads, row per thread, chunk loop)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_bench tools/tma_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cpa_arrive(uint32_t b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(b) : "memory"); }

constexpr int RING = 4;
// MODE 0..3 = A..D
template <int MODE>
__global__ void __launch_bounds__(192) k(const __grid_constant__ CUtensorMap m16, const __grid_constant__ CUtensorMap m128, const char* base,
                                          long long rows, int R, int C, int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const uint32_t slot_bytes = (uint32_t)C * ((R + 7) / 8 * 8) * 16;
    const uint32_t bars = smem_u32(sm) + RING * slot_bytes;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) { mbar_init(bars + 8 * i, MODE == 3 ? 129 : 1); mbar_init(bars + 8 * (RING + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    const long long tiles = rows / 128 - 2;
    if (tid < 128) {   // producers
        for (int it = 0; it < iters; ++it) {
            const int s = it % RING;
            if (it >= RING) mbar_wait(bars + 8 * (RING + s), ((it / RING) - 1) & 1);
            const long long tile = ((long long)blockIdx.x * iters + it) % tiles;
            const long long row0 = tile * 128 + (MODE == 1 ? 0 : 127);     // unaligned start where the view allows it
            const uint32_t dst = smem_u32(sm) + s * slot_bytes;
            const uint32_t full = bars + 8 * s;
            if (MODE == 0) {
                if (tid == 0) { mbar_expect(full, (uint32_t)C * R * 16); tma3d(dst, &m16, 0, (int)row0, 0, full); }
            } else if (MODE == 1) {
                if (tid == 0) { mbar_expect(full, (uint32_t)C * (R / 8 * 8) * 16); tma3d(dst, &m128, 0, (int)(row0 / 8), 0, full); }
            } else if (MODE == 2) {
                if (tid == 0) mbar_expect(full, (uint32_t)C * R * 16);
                if (tid < C) bulk1d(dst + tid * R * 16, base + ((long long)tid * rows + row0) * 16, (uint32_t)R * 16, full);
            } else {
                if (tid == 0) mbar_arrive(full);
                for (int m = tid; m < R; m += 128)
                    for (int c = 0; c < C; ++c) cpa16(dst + (c * R + m) * 16, base + ((long long)c * rows + row0 + m) * 16);
                cpa_arrive(full);
            }
        }
    } else if (tid == 128) {   // consumer
        for (int it = 0; it < iters; ++it) {
            const int s = it % RING;
            mbar_wait(bars + 8 * s, (it / RING) & 1);
            mbar_arrive(bars + 8 * (RING + s));
        }
    }
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = clock64() - t0;
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    const int NCH = 8;
    const long long rows = 524288;   // x 16 B x 8 chunks = 64 MB
    char* buf; cudaMalloc(&buf, rows * 16 * NCH); cudaMemset(buf, 0, rows * 16 * NCH);
    long long* cyc; cudaMalloc(&cyc, 8);
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int R = cfg == 0 ? 130 : 128, C = cfg == 0 ? 3 : 8;
        CUtensorMap m16, m128;
        { cuuint64_t d[3] = {8, (cuuint64_t)rows, NCH}, st[2] = {16, (cuuint64_t)rows * 16}; cuuint32_t b[3] = {8, (cuuint32_t)R, (cuuint32_t)C}, es[3] = {1, 1, 1};
          CUresult r = enc(&m16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r) printf("enc16 failed %d\n", (int)r); }
        { cuuint64_t d[3] = {64, (cuuint64_t)rows / 8, NCH}, st[2] = {128, (cuuint64_t)rows * 16}; cuuint32_t b[3] = {64, (cuuint32_t)(R / 8), (cuuint32_t)C}, es[3] = {1, 1, 1};
          CUresult r = enc(&m128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r) printf("enc128 failed %d\n", (int)r); }
        const int iters = 2000;
        const size_t smem = RING * (size_t)C * ((R + 7) / 8 * 8) * 16 + 256;
        const char* names[4] = {"A tensor 16B-inner (unaligned rows)", "B tensor 128B-inner (aligned rows)", "C 1-D bulk per chunk", "D cp.async 16 B per thread"};
        for (int mode = 0; mode < 4; ++mode) {
            auto launch = [&](int grid) {
                switch (mode) {
                    case 0: cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<grid, 192, smem>>>(m16, m128, buf, rows, R, C, iters, cyc); break;
                    case 1: cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<grid, 192, smem>>>(m16, m128, buf, rows, R, C, iters, cyc); break;
                    case 2: cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<2><<<grid, 192, smem>>>(m16, m128, buf, rows, R, C, iters, cyc); break;
                    default: cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<3><<<grid, 192, smem>>>(m16, m128, buf, rows, R, C, iters, cyc); break;
                }
            };
            for (int grid : {1, 148}) {
                launch(grid); cudaDeviceSynchronize();
                cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                cudaEventRecord(a); launch(grid); cudaEventRecord(b); cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, a, b);
                long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
                const double bytes = (double)C * R * 16 * iters;
                printf("R=%d C=%d %-38s grid %3d: %7.1f cycles/window  %6.1f B/cycle/SM  %7.1f GB/s total  (%s)\n", R, C, names[mode], grid, (double)c / iters,
                       bytes / (double)c, bytes * grid / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
