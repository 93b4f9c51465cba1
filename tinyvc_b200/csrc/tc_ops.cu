// Channels-last helper kernels of the tensor-core decoder path (layout changes, resampling,
// source assembly, output conv).  See tc_conv.cuh for the split-plane activation format.
#include "tc_conv.cuh"

namespace tvc {

namespace {

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

// Tiled transpose [B][C][T] (channels-first fp32) -> channels-last rows [B*T][cs]; MODE 0 = fp32 copy,
// MODE 1 = split planes.  Tile 32 channels x 32 time steps through shared memory, coalesced both ways.
template <int MODE>
__global__ void __launch_bounds__(256) cf_to_cl_kernel(const float* __restrict__ x, float* __restrict__ y32,
                                                       bf16* __restrict__ hi, bf16* __restrict__ lo, int C, int T,
                                                       int cs, int act, const float* __restrict__ extra0,
                                                       const float* __restrict__ extra1) {
    TVC_PDL_PROLOGUE();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
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
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + j, c = c0 + tx;
        if (t >= T || c >= cs) continue;
        const float v = tile[tx][j];
        const long long o = ((long long)b * T + t) * cs + c;
        if (MODE == 0) {
            y32[o] = v;
        } else {
            bf16 h, l;
            split_bf16(act_of(v, act), h, l);
            hi[o] = h;
            lo[o] = l;
        }
    }
}

// channels-last [B*T][cs] -> channels-first [B][C][T]; MODE 0 reads fp32, MODE 1 reads hi+lo planes.
template <int MODE>
__global__ void __launch_bounds__(256) cl_to_cf_kernel(const float* __restrict__ x32, const bf16* __restrict__ hi,
                                                       const bf16* __restrict__ lo, float* __restrict__ y, int C, int T,
                                                       int cs) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + j, c = c0 + tx;
        float v = 0.f;
        if (t < T && c < C) {
            const long long o = ((long long)b * T + t) * cs + c;
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

}  // namespace

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

}  // namespace tvc
