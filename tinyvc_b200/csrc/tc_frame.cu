// Channels-last kernels of the tensor-core decoder path that are not dense convs:
// linear resampling, ConvNeXt depth-wise conv + LayerNorm, GRN, and the k=7 output conv.
#include "tc_conv.cuh"
#include "tc_kernels.cuh"

namespace tvc {

namespace {

__device__ __forceinline__ void split_bf16(float v, bf16& h, bf16& l) {
    h = __float2bfloat16_rn(v);
    l = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(h)));
}
__device__ __forceinline__ uint32_t pack2(bf16 a, bf16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// store 4 consecutive channels of both planes (8 bytes each)
__device__ __forceinline__ void store_planes4(bf16* hi, bf16* lo, long long off, const float v[4]) {
    bf16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16(v[i], h[i], l[i]);
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
    *reinterpret_cast<uint2*>(lo + off) = make_uint2(pack2(l[0], l[1]), pack2(l[2], l[3]));
}

// store the 8 channels of one (row, chunk) of both planes (16 bytes each)
__device__ __forceinline__ void store_planes8(bf16* hi, bf16* lo, long long off, const float v[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bf16 h0, l0, h1, l1;
        split_bf16(v[2 * i], h0, l0);
        split_bf16(v[2 * i + 1], h1, l1);
        h[i] = pack2(h0, h1);
        l[i] = pack2(l0, l1);
    }
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------------
// interp_cl:  F.interpolate(x, scale_factor=..., mode='linear') along time for chunk-major channels-last
// fp32 input with B*Tin rows (decoder.py:148 Downsample, :174 Upsample), exact ATen coordinate
// arithmetic (tvc_common.cuh lin_coord/lin_blend).  Writes any of: fp32, raw split planes,
// leaky_relu(0.1) split planes (the convs that follow apply leaky_relu first), all with B*Tout rows.
// One thread per (8-channel chunk, output row): blockIdx.y = chunk, consecutive threads = consecutive rows of it.
// ---------------------------------------------------------------------------------------------
// Windowed form (output pruning, nets_tc.cu): the output tensor holds rows [out_off, out_off + Tout) of every utterance's
// full-length result and the input tensor rows [in_off, in_off + Tin_c) of its full-length input (Tin rows); coordinates are
// evaluated on the full lengths, so the values are those of the full-length call.
__global__ void __launch_bounds__(256) interp_cl_kernel(const float* __restrict__ x, int Tin, int Tout, float scale,
                                                        long long rows_in, long long rows_out, float* __restrict__ y32,
                                                        bf16* __restrict__ r_hi, bf16* __restrict__ r_lo,
                                                        bf16* __restrict__ a_hi, bf16* __restrict__ a_lo, int a_pad,
                                                        int Tin_c, int in_off, int out_off) {
    TVC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_out) return;
    const long long q = blockIdx.y;              // chunk
    const unsigned bu = (unsigned)row / (unsigned)Tout;      // rows < 2^31 (checked at the API)
    const long long b = bu;
    const int t = (int)((unsigned)row - bu * (unsigned)Tout);
    const LinCoord c = lin_coord(t + out_off, scale, Tin);
    const float4* p0 = reinterpret_cast<const float4*>(x + (q * rows_in + b * Tin_c + (c.i0 - in_off)) * 8);
    const float4* p1 = reinterpret_cast<const float4*>(x + (q * rows_in + b * Tin_c + (c.i1 - in_off)) * 8);
    const float4 x0 = __ldg(p0), x1 = __ldg(p0 + 1), z0 = __ldg(p1), z1 = __ldg(p1 + 1);
    float v[8] = {lin_blend(x0.x, z0.x, c), lin_blend(x0.y, z0.y, c), lin_blend(x0.z, z0.z, c), lin_blend(x0.w, z0.w, c),
                  lin_blend(x1.x, z1.x, c), lin_blend(x1.y, z1.y, c), lin_blend(x1.z, z1.z, c), lin_blend(x1.w, z1.w, c)};
    const long long o = (q * rows_out + row) * 8;
    if (y32) {
        reinterpret_cast<float4*>(y32 + o)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(y32 + o)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (r_hi) store_planes8(r_hi, r_lo, o, v);
    if (a_hi) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = leaky01(v[k]);
        if (a_pad == 0) {
            store_planes8(a_hi, a_lo, o, v);
        } else {
            // stored replicate padding (tc_conv.cuh, padded mode): a_pad extra rows on either side of every utterance
            const long long Tp = Tout + 2 * a_pad, rows_p = (rows_out / Tout) * Tp;
            const long long op = (q * rows_p + b * Tp + a_pad + t) * 8;
            store_planes8(a_hi, a_lo, op, v);
            if (t == 0)
                for (int k = 1; k <= a_pad; ++k) store_planes8(a_hi, a_lo, op - 8 * k, v);
            if (t == Tout - 1)
                for (int k = 1; k <= a_pad; ++k) store_planes8(a_hi, a_lo, op + 8 * k, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dwconv_ln_cl (C = 128): depth-wise k=7 conv (dilation 1, replicate padding) + LayerNorm over
// channels (convnext.py:42-43,52-53) on chunk-major channels-last fp32 -> split planes.  One warp per
// row, 4 channels per lane (half a chunk); the LayerNorm statistics are two warp-shuffle reductions.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv_ln_cl_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ w7,   // [7][128] repacked
                                                           const float* __restrict__ wb, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, bf16* __restrict__ hi,
                                                           bf16* __restrict__ lo, int T, long long rows) {
    TVC_PDL_PROLOGUE();
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long b = row / T;
    const int t = (int)(row - b * T);
    const float4 bv = __ldg(reinterpret_cast<const float4*>(wb) + lane);
    float v[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        int tt = t + j - 3;
        tt = tt < 0 ? 0 : (tt > T - 1 ? T - 1 : tt);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + cm(b * T + tt, lane * 4, rows)));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w7 + j * 128) + lane);
        v[0] = fmaf(wv.x, xv.x, v[0]); v[1] = fmaf(wv.y, xv.y, v[1]);
        v[2] = fmaf(wv.z, xv.z, v[2]); v[3] = fmaf(wv.w, xv.w, v[3]);
    }
    float s = (v[0] + v[1]) + (v[2] + v[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 128.0f);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float d = v[k] - mean;
        q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + 1e-5f);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
    float o4[4] = {fmaf((v[0] - mean) * rstd, g.x, be.x), fmaf((v[1] - mean) * rstd, g.y, be.y),
                   fmaf((v[2] - mean) * rstd, g.z, be.z), fmaf((v[3] - mean) * rstd, g.w, be.w)};
    store_planes4(hi, lo, cm(row, lane * 4, rows), o4);
}

// ---------------------------------------------------------------------------------------------
// cnxt_ln_cl<C>: the general form for the Encoder's stacks (encoder.py:27-31,84-88; convnext.py:42-43,52-53):
// C = 128 or 384 channels, depth-wise k = 7 conv with dilation `dil` (replicate padding = clamped time index) followed by
// LayerNorm over channels, or LayerNorm alone (w7 == nullptr: the stacks' input norm).  One warp per row; lane l owns the
// float4 groups l, l + 32, ... of the row (4 channels each).  Outputs: split planes (the next conv's operand) and / or fp32.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) cnxt_ln_cl_kernel(const float* __restrict__ x, const float* __restrict__ w7,   // [7][C] repacked
                                                         const float* __restrict__ wb, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                                         float* __restrict__ y32, int T, int dil, long long rows) {
    TVC_PDL_PROLOGUE();
    constexpr int G = C / 128;                    // float4 groups per lane
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long b = row / T;
    const int t = (int)(row - b * T);
    float v[G][4];
    if (w7) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(wb) + lane + 32 * g);
            v[g][0] = bv.x; v[g][1] = bv.y; v[g][2] = bv.z; v[g][3] = bv.w;
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            int tt = t + (j - 3) * dil;
            tt = tt < 0 ? 0 : (tt > T - 1 ? T - 1 : tt);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int c4 = lane + 32 * g;
                const float4 xv = __ldg(reinterpret_cast<const float4*>(x + cm(b * T + tt, c4 * 4, rows)));
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w7 + j * C) + c4);
                v[g][0] = fmaf(wv.x, xv.x, v[g][0]); v[g][1] = fmaf(wv.y, xv.y, v[g][1]);
                v[g][2] = fmaf(wv.z, xv.z, v[g][2]); v[g][3] = fmaf(wv.w, xv.w, v[g][3]);
            }
        }
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + cm(row, (lane + 32 * g) * 4, rows)));
            v[g][0] = xv.x; v[g][1] = xv.y; v[g][2] = xv.z; v[g][3] = xv.w;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) s += (v[g][0] + v[g][1]) + (v[g][2] + v[g][3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / (float)C);
    float q = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d = v[g][k] - mean;
            q = fmaf(d, d, q);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / (float)C) + 1e-5f);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int c4 = lane + 32 * g;
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        float o4[4] = {fmaf((v[g][0] - mean) * rstd, ga.x, be.x), fmaf((v[g][1] - mean) * rstd, ga.y, be.y),
                       fmaf((v[g][2] - mean) * rstd, ga.z, be.z), fmaf((v[g][3] - mean) * rstd, ga.w, be.w)};
        const long long o = cm(row, c4 * 4, rows);
        if (hi) store_planes4(hi, lo, o, o4);
        if (y32) *reinterpret_cast<float4*>(y32 + o) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// grn_apply_wide_cl: GRN (convnext.py:23-34) for any channel count (the Encoder's 768 / 256): one block per utterance,
// a thread owns channels tid, tid + 256, ...; fixed summation order (deterministic).
// ---------------------------------------------------------------------------------------------
// One thread per channel (up to 1024 channels): a thread's two passes over its T rows are short chains of independent loads.
__global__ void __launch_bounds__(1024) grn_apply_wide_cl_kernel(const float* __restrict__ y, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, bf16* __restrict__ hi,
                                                                 bf16* __restrict__ lo, int C, int T) {
    TVC_PDL_PROLOGUE();
    __shared__ float part[32];
    const int b = blockIdx.x, c = threadIdx.x, lane = c & 31, warp = c >> 5;
    const long long R = (long long)gridDim.x * T;
    const long long base = cm((long long)b * T, c < C ? c : 0, R);
    float s = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int t = 0; t < T; ++t) {
            const float v = __ldg(y + base + (long long)t * 8);
            s = fmaf(v, v, s);
        }
    }
    const float g = sqrtf(s);
    // the block total is formed in the order of the 256-thread version (thread tid summed its channels tid, tid + 256, tid + 512,
    // then a shuffle tree per warp, then the eight warp sums in order), so the statistic keeps its bits
    __shared__ float gs[1024];
    gs[c] = c < C ? g : 0.f;
    __syncthreads();
    if (c < 256) {
        float tot = 0.f;
        for (int cc = c; cc < C; cc += 256) tot += gs[cc];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) part[warp] = tot;
    }
    __syncthreads();
    float all = 0.f;
    for (int i = 0; i < 8; ++i) all += part[i];
    if (c >= C) return;
    const float denom = all / (float)C + 1e-6f;
    const float scale = fmaf(__ldg(gamma + c), g / denom, 1.0f);
    const float bt = __ldg(beta + c);
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
        const long long o = base + (long long)t * 8;
        bf16 h, l;
        split_bf16(fmaf(__ldg(y + o), scale, bt), h, l);
        hi[o] = h;
        lo[o] = l;
    }
}

// ---------------------------------------------------------------------------------------------
// grn_apply_cl: GRN (convnext.py:23-34) on chunk-major channels-last fp32 (B*T rows, C channels) -> split planes:
//   g[c] = sqrt(sum_t y^2),  n = g / (mean_c g + 1e-6),  out = y * (gamma*n + 1) + beta
// One block per utterance, thread = channel (coalesced over channels), fixed summation order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grn_apply_cl_kernel(const float* __restrict__ y, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, bf16* __restrict__ hi,
                                                           bf16* __restrict__ lo, int C, int T) {
    TVC_PDL_PROLOGUE();
    __shared__ float part[8];
    const int b = blockIdx.x, c = threadIdx.x, lane = c & 31, warp = c >> 5;
    const long long R = (long long)gridDim.x * T;
    const long long base = cm((long long)b * T, c, R);        // (row b*T, channel c); consecutive rows are 8 floats apart
    float s = 0.f;
    if (c < C)
        for (int t = 0; t < T; ++t) {
            const float v = __ldg(y + base + (long long)t * 8);
            s = fmaf(v, v, s);
        }
    const float g = sqrtf(s);
    float tot = c < C ? g : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) part[warp] = tot;
    __syncthreads();
    float all = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) all += part[i];
    if (c >= C) return;
    const float scale = fmaf(__ldg(gamma + c), g / (all / (float)C + 1e-6f), 1.0f);
    const float bt = __ldg(beta + c);
    for (int t = 0; t < T; ++t) {
        const long long o = base + (long long)t * 8;
        bf16 h, l;
        split_bf16(fmaf(__ldg(y + o), scale, bt), h, l);
        hi[o] = h;
        lo[o] = l;
    }
}

// ---------------------------------------------------------------------------------------------
// out_conv_k7_cl: FilterNet.output_layer, Conv1d(24 -> 1, k = 7, replicate pad 3) (decoder.py:220,233)
// on chunk-major channels-last fp32 (B*T rows, 24 channels = 3 chunks) -> waveform [B][T].  One thread per output
// sample; a warp reads 32 consecutive rows of a chunk = 1 KB contiguous, neighbouring taps hit L1.
// (A shared-memory staged variant with 4 outputs per thread measured slower: 58 vs 35 us at config 2.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) out_conv_k7_cl_kernel(const float* __restrict__ x, const float* __restrict__ w,   // [24][7] torch layout
                                                             const float* __restrict__ bias, float* __restrict__ y, int T,
                                                             long long rows) {
    TVC_PDL_PROLOGUE();
    __shared__ float ws[7][24];
    for (int e = threadIdx.x; e < 24 * 7; e += blockDim.x) ws[e % 7][e / 7] = __ldg(w + e);
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const long long b = row / T;
    const int t = (int)(row - b * T);
    float acc = __ldg(bias);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        int tt = t + j - 3;
        tt = tt < 0 ? 0 : (tt > T - 1 ? T - 1 : tt);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + cm(b * T + tt, q * 4, rows)));
            acc = fmaf(ws[j][q * 4 + 0], v.x, acc);
            acc = fmaf(ws[j][q * 4 + 1], v.y, acc);
            acc = fmaf(ws[j][q * 4 + 2], v.z, acc);
            acc = fmaf(ws[j][q * 4 + 3], v.w, acc);
        }
    }
    y[row] = acc;
}

}  // namespace

int interp_cl(const float* x, int B, int Tin, int Tout, float scale, int C, float* y32, bf16* r_hi, bf16* r_lo, bf16* a_hi,
              bf16* a_lo, cudaStream_t s, int a_pad, int Tin_c, int in_off, int out_off) {
    TVC_REQUIRE(C % 8 == 0, "interp_cl: channel count %d must be a multiple of 8", C);
    if (Tin_c <= 0) Tin_c = Tin;
    TVC_REQUIRE((Tin_c == Tin && in_off == 0 && out_off == 0) || a_pad == 0, "interp_cl: the windowed form has no stored padding");
    const long long rows_in = (long long)B * Tin_c, rows_out = (long long)B * Tout;
    TVC_REQUIRE(rows_out < (1LL << 31) && rows_in < (1LL << 31), "interp_cl: too many rows");
    TVC_LAUNCH_PDL(interp_cl_kernel, dim3(cdiv(rows_out, 256), C / 8), 256, 0, s, x, Tin, Tout, scale, rows_in, rows_out, y32, r_hi, r_lo, a_hi, a_lo, a_pad,
                   Tin_c, in_off, out_off);
    TVC_LAUNCH_CHECK();
    return 0;
}

// rows [t0, t0 + Tc) of every utterance of a chunk-major plane pair -> a compact pair with Tc rows per utterance
__global__ void __launch_bounds__(256) slice_planes_cl_kernel(const bf16* __restrict__ s_hi, const bf16* __restrict__ s_lo,
                                                              bf16* __restrict__ d_hi, bf16* __restrict__ d_lo, int T, int Tc, int t0,
                                                              long long rows_in, long long rows_out) {
    TVC_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_out) return;
    const long long q = blockIdx.y;
    const long long b = row / Tc;
    const int r = (int)(row - b * Tc);
    const long long si = (q * rows_in + b * T + t0 + r) * 8, di = (q * rows_out + row) * 8;
    *reinterpret_cast<uint4*>(d_hi + di) = __ldg(reinterpret_cast<const uint4*>(s_hi + si));
    *reinterpret_cast<uint4*>(d_lo + di) = __ldg(reinterpret_cast<const uint4*>(s_lo + si));
}
int slice_planes_cl(const bf16* s_hi, const bf16* s_lo, bf16* d_hi, bf16* d_lo, int B, int T, int C, int t0, int Tc, cudaStream_t s) {
    TVC_REQUIRE(C % 8 == 0 && t0 >= 0 && Tc > 0 && t0 + Tc <= T, "slice_planes_cl: rows [%d, %d) of %d, %d channels", t0, t0 + Tc, T, C);
    const long long rows_in = (long long)B * T, rows_out = (long long)B * Tc;
    TVC_LAUNCH_PDL(slice_planes_cl_kernel, dim3(cdiv(rows_out, 256), C / 8), 256, 0, s, s_hi, s_lo, d_hi, d_lo, T, Tc, t0, rows_in, rows_out);
    TVC_LAUNCH_CHECK();
    return 0;
}

int dwconv_ln_cl(const float* x, const float* w7, const float* wb, const float* gamma, const float* beta,
                 bf16* hi, bf16* lo, int B, int T, cudaStream_t s) {
    const long long rows = (long long)B * T;
    TVC_LAUNCH_PDL(dwconv_ln_cl_kernel, cdiv(rows * 32, 256), 256, 0, s, x, w7, wb, gamma, beta, hi, lo, T, rows);
    TVC_LAUNCH_CHECK();
    return 0;
}

int cnxt_ln_cl(const float* x, const float* w7, const float* wb, const float* gamma, const float* beta, bf16* hi, bf16* lo,
               float* y32, int B, int C, int T, int dil, cudaStream_t s) {
    const long long rows = (long long)B * T;
    TVC_REQUIRE(C == 128 || C == 384, "cnxt_ln_cl: C=%d (128 or 384)", C);
    if (C == 128) TVC_LAUNCH_PDL(cnxt_ln_cl_kernel<128>, cdiv(rows * 32, 256), 256, 0, s, x, w7, wb, gamma, beta, hi, lo, y32, T, dil, rows);
    else TVC_LAUNCH_PDL(cnxt_ln_cl_kernel<384>, cdiv(rows * 32, 256), 256, 0, s, x, w7, wb, gamma, beta, hi, lo, y32, T, dil, rows);
    TVC_LAUNCH_CHECK();
    return 0;
}

int grn_apply_cl(const float* y, const float* gamma, const float* beta, bf16* hi, bf16* lo, int B, int C, int T,
                 cudaStream_t s) {
    if (C > 256) {
        TVC_REQUIRE(C <= 1024, "grn_apply_cl: C=%d > 1024", C);
        TVC_LAUNCH_PDL(grn_apply_wide_cl_kernel, B, (C + 31) / 32 * 32, 0, s, y, gamma, beta, hi, lo, C, T);
        TVC_LAUNCH_CHECK();
        return 0;
    }
    TVC_LAUNCH_PDL(grn_apply_cl_kernel, B, 256, 0, s, y, gamma, beta, hi, lo, C, T);
    TVC_LAUNCH_CHECK();
    return 0;
}

int out_conv_k7_cl(const float* x, const float* w, const float* bias, float* y, int B, int T, cudaStream_t s) {
    const long long rows = (long long)B * T;
    TVC_LAUNCH_PDL(out_conv_k7_cl_kernel, cdiv(rows, 256), 256, 0, s, x, w, bias, y, T, rows);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
