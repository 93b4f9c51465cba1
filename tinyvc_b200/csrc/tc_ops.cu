// Channels-last helper kernels of the tensor-core decoder path (layout changes, resampling,
// source assembly, output conv).  See tc_conv.cuh for the split-plane activation format.
#include "tc_conv.cuh"

namespace tvc {

namespace {

// FP32 FMA peak probe: 8 independent chains per thread, all operands in registers.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;       // never true for the arguments used; keeps the chains alive
}

__device__ __forceinline__ float act_of(float v, int act) {
    switch (act) {
        case TC_ACT_LRELU: return leaky01(v);
        case TC_ACT_GELU: return gelu_erf(v);
        case TC_ACT_ELU1: return elu_plus1(v);
        default: return v;
    }
}
__device__ __forceinline__ void split_bf16(float v, bf16& h, bf16& l) {
    h = __float2bfloat16_rn(v);
    l = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(h)));
}

// Tiled transpose [B][C][T] (channels-first fp32) -> chunk-major channels-last (tc_conv.cuh) with `cs` channels of
// capacity; MODE 0 = fp32 copy, MODE 1 = split planes.  Tile 32 channels x 32 time steps through shared memory;
// the write side handles one (time step, 8-channel chunk) per thread = one 16 / 32-byte row of a chunk array.
template <int MODE>
__global__ void __launch_bounds__(256) cf_to_cl_kernel(const float* __restrict__ x, float* __restrict__ y32,
                                                       bf16* __restrict__ hi, bf16* __restrict__ lo, int C, int T,
                                                       int cs, int act, const float* __restrict__ extra0,
                                                       const float* __restrict__ extra1) {
    TVC_PDL_PROLOGUE();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long R = (long long)gridDim.z * T;
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, t = t0 + tx;
        float v = 0.f;
        if (t < T) {
            if (c < C) v = __ldg(x + ((long long)b * C + c) * T + t);
            else if (c == C && extra0) v = __ldg(extra0 + (long long)b * T + t);       // per-row scalars appended as
            else if (c == C + 1 && extra1) v = __ldg(extra1 + (long long)b * T + t);   // channels C, C+1
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int q = threadIdx.x >> 5, t = t0 + tx;          // chunk q of this tile, time step t
        const int c = c0 + q * 8;
        if (t < T && c < cs) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = tile[q * 8 + e][tx];
            const long long o = cm((long long)b * T + t, c, R);
            if (MODE == 0) {
                reinterpret_cast<float4*>(y32 + o)[0] = make_float4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<float4*>(y32 + o)[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bf16 h0, l0, h1, l1;
                    split_bf16(act_of(v[2 * e], act), h0, l0);
                    split_bf16(act_of(v[2 * e + 1], act), h1, l1);
                    h[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    l[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
            }
        }
    }
}

// chunk-major channels-last -> channels-first [B][C][T]; MODE 0 reads fp32, MODE 1 reads hi+lo planes (parity probes).
template <int MODE>
__global__ void __launch_bounds__(256) cl_to_cf_kernel(const float* __restrict__ x32, const bf16* __restrict__ hi,
                                                       const bf16* __restrict__ lo, float* __restrict__ y, int C, int T,
                                                       int cs) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long R = (long long)gridDim.z * T;
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + j, c = c0 + tx;
        float v = 0.f;
        if (t < T && c < C) {
            const long long o = cm((long long)b * T + t, c, R);
            v = MODE == 0 ? __ldg(x32 + o) : __fadd_rn(__bfloat162float(hi[o]), __bfloat162float(lo[o]));
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, t = t0 + tx;
        if (c < C && t < T) y[((long long)b * C + c) * T + t] = tile[tx][j];
    }
}

__global__ void __launch_bounds__(256) plane_repad_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T, int pad_in,
                                                         int pad_out, long long rows_in, long long rows_out) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_out) return;
    const long long q = blockIdx.y, Tpo = T + 2 * pad_out, Tpi = T + 2 * pad_in;
    const long long b = row / Tpo;
    int t = (int)(row - b * Tpo) - pad_out;
    t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
    out[q * rows_out + row] = in[q * rows_in + b * Tpi + pad_in + t];
}

}  // namespace

int plane_repad(const bf16* in, bf16* out, int B, int T, int cs, int pad_in, int pad_out, cudaStream_t s) {
    const long long rows_in = (long long)B * (T + 2 * pad_in), rows_out = (long long)B * (T + 2 * pad_out);
    plane_repad_kernel<<<dim3(cdiv(rows_out, 256), cs / 8), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out),
                                                                         T, pad_in, pad_out, rows_in, rows_out);
    TVC_LAUNCH_CHECK();
    return 0;
}

int cf_to_planes(const float* x, bf16* hi, bf16* lo, int B, int C, int T, int cs, int act, cudaStream_t s,
                 const float* extra0, const float* extra1) {
    dim3 grid(cdiv(T, 32), cdiv(cs, 32), B);
    TVC_LAUNCH_PDL(cf_to_cl_kernel<1>, grid, 256, 0, s, x, nullptr, hi, lo, C, T, cs, act, extra0, extra1);
    TVC_LAUNCH_CHECK();
    return 0;
}
int cf_to_cl(const float* x, float* y, int B, int C, int T, int cs, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), cdiv(cs, 32), B);
    cf_to_cl_kernel<0><<<grid, 256, 0, s>>>(x, y, nullptr, nullptr, C, T, cs, 0, nullptr, nullptr);
    TVC_LAUNCH_CHECK();
    return 0;
}
int cl_to_cf(const float* x, float* y, int B, int C, int T, int cs, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), cdiv(C, 32), B);
    cl_to_cf_kernel<0><<<grid, 256, 0, s>>>(x, nullptr, nullptr, y, C, T, cs);
    TVC_LAUNCH_CHECK();
    return 0;
}
int planes_to_cf(const bf16* hi, const bf16* lo, float* y, int B, int C, int T, int cs, cudaStream_t s) {
    dim3 grid(cdiv(T, 32), cdiv(C, 32), B);
    cl_to_cf_kernel<1><<<grid, 256, 0, s>>>(nullptr, hi, lo, y, C, T, cs);
    TVC_LAUNCH_CHECK();
    return 0;
}

int measure_fp32_peak(double* tflops, cudaStream_t s) {
    int dev = 0, sms = 0;
    TVC_CUDA(cudaGetDevice(&dev));
    TVC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* sink = nullptr;
    TVC_CUDA(cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t a, b;
    TVC_CUDA(cudaEventCreate(&a));
    TVC_CUDA(cudaEventCreate(&b));
    const int iters = 1 << 14, blocks = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {            // first repetition warms the clocks up
        cudaEventRecord(a, s);
        fma_peak_kernel<<<blocks, 256, 0, s>>>(sink, iters, 0.999f, 0.001f);
        cudaEventRecord(b, s);
        TVC_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double fl = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
        if (ms > 0.f && fl / (ms * 1e-3) / 1e12 > best) best = fl / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    TVC_CUDA(cudaGetLastError());
    *tflops = best;
    return 0;
}

}  // namespace tvc
