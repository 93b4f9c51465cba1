// SOLA alignment + cross-fade of StreamInfer.audio_callback (module/infer/stream.py:74-95), batched
// over S independent streams and kept on the device (the reference syncs with .item() every tick).
//
// per stream:  temp = y[-(block+cross+search+delay) : -delay]
//   nom[i] = sum_{j<cross} temp[i+j] * sola[j]            F.conv1d(conv_input, sola_buffer)      (:77)
//   den[i] = sqrt(sum_{j<cross} temp[i+j]^2 + 1e-8)       F.conv1d(conv_input**2, ones)          (:78)
//   shift  = argmax_i nom[i]/den[i],  i in [0, search]                                           (:79)
//   out    = temp[shift : shift+block+cross];  out[:cross] = out[:cross]*fade_in + sola*fade_out  (:80,90-92)
//   sola'  = out[-cross:] ; return out[:block]                                                   (:94-95)
#include "tvc_kernels.cuh"

namespace tvc {

__global__ void __launch_bounds__(1024) sola_kernel(const float* __restrict__ y, int y_len, float* __restrict__ sola_buf,
                                                    const float* __restrict__ fade_in, float* __restrict__ out_block,
                                                    int* __restrict__ shift_out, int block, int cross, int search,
                                                    int delay) {
    extern __shared__ float sm[];
    float* temp = sm;                           // [block+cross+search]
    float* sola = temp + block + cross + search;   // [cross]
    __shared__ float best_v[32];
    __shared__ int best_i[32];
    __shared__ int s_shift;
    const int s = blockIdx.x;
    const int tl = block + cross + search;
    const float* ys = y + (long long)s * y_len + (y_len - tl - delay);
    for (int i = threadIdx.x; i < tl; i += blockDim.x) temp[i] = ys[i];
    for (int i = threadIdx.x; i < cross; i += blockDim.x) sola[i] = sola_buf[(long long)s * cross + i];
    __syncthreads();
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i <= search; i += blockDim.x) {
        float nom = 0.f, den = 0.f;
        for (int j = 0; j < cross; ++j) {
            const float v = temp[i + j];
            nom = fmaf(v, sola[j], nom);
            den = fmaf(v, v, den);
        }
        const float r = __fdiv_rn(nom, sqrtf(__fadd_rn(den, 1e-8f)));
        if (r > bv || (r == bv && i < bi)) {
            bv = r;
            bi = i;
        }
    }
    // block arg-max (first index on ties, like torch.argmax)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        best_v[warp] = bv;
        best_i[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = best_v[0];
        int ix = best_i[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (best_v[w] > v || (best_v[w] == v && best_i[w] < ix)) {
                v = best_v[w];
                ix = best_i[w];
            }
        if (ix == 0x7fffffff) ix = 0;   // all-NaN correlation: torch.argmax would return a NaN position; use 0
        s_shift = ix;
        shift_out[s] = ix;
    }
    __syncthreads();
    const int sh = s_shift;
    for (int i = threadIdx.x; i < block; i += blockDim.x) {
        float v = temp[sh + i];
        if (i < cross) {
            const float fi = __ldg(fade_in + i);
            v = __fadd_rn(__fmul_rn(v, fi), __fmul_rn(sola[i], __fsub_rn(1.0f, fi)));
        }
        out_block[(long long)s * block + i] = v;
    }
    __syncthreads();   // every read of the old sola buffer is done
    for (int i = threadIdx.x; i < cross; i += blockDim.x) {
        const int p = block + i;   // position inside out = temp[sh:]
        float v = temp[sh + p];
        if (p < cross) {           // only when block < cross: the tail overlaps the faded head
            const float fi = __ldg(fade_in + p);
            v = __fadd_rn(__fmul_rn(v, fi), __fmul_rn(sola[p], __fsub_rn(1.0f, fi)));
        }
        sola_buf[(long long)s * cross + i] = v;
    }
}

int sola_run(const float* y, int y_len, float* sola_buf, const float* fade_in, float* out_block, int* shift_out, int S,
             int block, int cross, int search, int delay, cudaStream_t s) {
    TVC_REQUIRE(y_len >= block + cross + search + delay, "sola: window of %d samples is shorter than block+cross+search+delay = %d",
                y_len, block + cross + search + delay);
    const size_t smem = sizeof(float) * (size_t)(block + 2 * cross + search);
    TVC_REQUIRE(smem <= 200 * 1024, "sola: block/cross/search too large for shared memory (%zu bytes)", smem);
    static bool attr_done = false;
    if (!attr_done) {
        TVC_CUDA(cudaFuncSetAttribute(sola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    sola_kernel<<<S, 1024, smem, s>>>(y, y_len, sola_buf, fade_in, out_block, shift_out, block, cross, search, delay);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
