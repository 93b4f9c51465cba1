// Frame-rate and element-wise kernels of the TinyVC hot path (everything that is not a dense conv).
#include "tvc_kernels.cuh"

namespace tvc {

// ---------------------------------------------------------------------------------------------
// frame_prep:  e_fr[b,t] = max_{i<480} energy[b,480t+i]      (decoder.py:127  F.max_pool1d(energy,480,480))
//              lf0[b,t]  = log(relu(f0[b,t]) + 1e-6)          (decoder.py:128,223)
// one warp per frame, coalesced 128 B reads.
// ---------------------------------------------------------------------------------------------
__global__ void frame_prep_kernel(const float* __restrict__ energy, const float* __restrict__ f0,
                                  float* __restrict__ e_fr, float* __restrict__ lf0, long long nframes) {
    TVC_PDL_PROLOGUE();
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nframes) return;
    const float* e = energy + w * kFrame;
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kFrame / 32; ++i) m = fmaxf(m, __ldg(e + i * 32 + lane));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        e_fr[w] = m;
        lf0[w] = logf(fmaxf(__ldg(f0 + w), 0.f) + 1e-6f);
    }
}

int frame_prep(const float* energy, const float* f0, float* e_fr, float* lf0, int B, int Lf, cudaStream_t s) {
    const long long nf = (long long)B * Lf;
    const int threads = 256;
    TVC_LAUNCH_PDL(frame_prep_kernel, cdiv(nf * 32, threads), threads, 0, s, energy, f0, e_fr, lf0, nf);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// rank1_add:  x[b,c,t] = (x + (w1[c]*a1[b,t] + b1[c])) + (w2[c]*a2[b,t] + b2[c])
// The 1->C 1x1 convs `energy_in` / `f0_in` (decoder.py:120-121,202) are outer products; the
// reference adds them in this order (decoder.py:128,223).  a1 may be null (FilterNet has no energy_in).
// ---------------------------------------------------------------------------------------------
__global__ void rank1_add_kernel(float* __restrict__ x, const float* __restrict__ a1, const float* __restrict__ w1,
                                 const float* __restrict__ b1, const float* __restrict__ a2,
                                 const float* __restrict__ w2, const float* __restrict__ b2, int C, int T,
                                 long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    float v = x[i];
    if (a1) v = __fadd_rn(v, fmaf(__ldg(w1 + c), __ldg(a1 + b * T + t), __ldg(b1 + c)));
    v = __fadd_rn(v, fmaf(__ldg(w2 + c), __ldg(a2 + b * T + t), __ldg(b2 + c)));
    x[i] = v;
}

int rank1_add(float* x, const float* a1, const float* w1, const float* b1, const float* a2, const float* w2,
              const float* b2, int B, int C, int T, cudaStream_t s) {
    const long long total = (long long)B * C * T;
    rank1_add_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, a1, w1, b1, a2, w2, b2, C, T, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// dwconv_ln:  y = LayerNorm_C( depthwise_conv_k7(x, dilation, replicate pad) )       (convnext.py:42-43,52-53)
// or, with w == nullptr, just the channel LayerNorm (encoder.py:28,35 / convnext.py:17-19).
// Block = one utterance x 32 consecutive frames x all C channels; lane = frame (coalesced),
// warp strides over channels.  The conv result is parked in shared memory ([C][33], conflict
// free in both directions) for the two-pass mean / biased variance over channels.
//   y = ((v - mean) * rstd) * gamma + beta     with rstd = 1/sqrt(var + eps), eps = 1e-5
// In LN-only mode the kernel is safe in place (a block only touches its own tile).
// ---------------------------------------------------------------------------------------------
constexpr int kLnTT = 32;
__global__ void __launch_bounds__(256) dwconv_ln_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ w, const float* __restrict__ wb,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int C, int T, int dil, float eps) {
    extern __shared__ float sm[];
    float* tile = sm;                       // [C][33]
    float* s_mean = sm + (size_t)C * 33;    // [32]
    float* s_rstd = s_mean + 32;            // [32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int b = blockIdx.y;
    const int t = blockIdx.x * kLnTT + lane;
    const bool ok = t < T;
    const float* xb = x + (long long)b * C * T;

    int tt[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        int q = t + (j - 3) * dil;
        tt[j] = q < 0 ? 0 : (q > T - 1 ? T - 1 : q);
    }
    // four channels per step with every load issued before the first use (frames past the end read a clamped, legal
    // address and are masked afterwards): a short utterance batch is a chain of load round trips otherwise
    const int tl = ok ? t : T - 1;
    for (int c0 = warp; c0 < C; c0 += 4 * nwarp) {
        float xv[4][7], wv[4][7], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = min(c0 + u * nwarp, C - 1);
            const float* xc = xb + (long long)c * T;
            if (w) {
                bv[u] = __ldg(wb + c);
#pragma unroll
                for (int j = 0; j < 7; ++j) { wv[u][j] = __ldg(w + c * 7 + j); xv[u][j] = __ldg(xc + tt[j]); }
            } else {
                bv[u] = xc[tl];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * nwarp;
            if (c >= C) break;
            float v = bv[u];
            if (w) {
#pragma unroll
                for (int j = 0; j < 7; ++j) v = fmaf(wv[u][j], xv[u][j], v);
            }
            tile[c * 33 + lane] = ok ? v : 0.f;
        }
    }
    __syncthreads();
    // per-frame statistics: warp `warp` handles frames warp, warp+nwarp, ...
    for (int col = warp; col < kLnTT; col += nwarp) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += tile[c * 33 + col];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)C;
        float q = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float d = tile[c * 33 + col] - mean;
            q = fmaf(d, d, q);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) {
            s_mean[col] = mean;
            s_rstd[col] = 1.0f / sqrtf(q / (float)C + eps);
        }
    }
    __syncthreads();
    if (!ok) return;
    const float mean = s_mean[lane], rstd = s_rstd[lane];
    float* yb = y + (long long)b * C * T;
    for (int c = warp; c < C; c += nwarp) {
        const float v = tile[c * 33 + lane];
        yb[(long long)c * T + t] = fmaf((v - mean) * rstd, __ldg(gamma + c), __ldg(beta + c));
    }
}

int dwconv_ln(const float* x, float* y, const float* w, const float* wb, const float* gamma, const float* beta,
              int B, int C, int T, int dil, cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)C * 33 + 64);
    TVC_REQUIRE(smem <= 200 * 1024, "dwconv_ln: C=%d too large for the shared-memory tile", C);
    TVC_REQUIRE(!(w && x == y), "dwconv_ln: the conv form is not in-place safe");
    static PerDeviceOnce attr;
    TVC_TRY(attr.run([] { TVC_CUDA(cudaFuncSetAttribute(dwconv_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); return 0; }));
    dim3 grid(cdiv(T, kLnTT), B);
    dwconv_ln_kernel<<<grid, 256, smem, s>>>(x, y, w, wb, gamma, beta, C, T, dil, 1e-5f);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// grn_scale:  g[b,c] = sqrt(sum_t y[b,c,t]^2);  n = g / (mean_c g + 1e-6);  scale[b,c] = gamma[c]*n + 1
// so that GRN(y) = gamma*(y*n) + beta + y = y*scale + beta   (convnext.py:31-34) can be applied by the
// consuming 1x1 conv's PRE_AFFINE prologue.  One block per utterance; fixed reduction order
// (no atomics) so results do not depend on batch composition or rank count.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grn_scale_kernel(const float* __restrict__ y, const float* __restrict__ gamma,
                                                        float* __restrict__ scale, int C, int T) {
    extern __shared__ float g[];   // [C] + [8]
    float* part = g + C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int b = blockIdx.x;
    const float* yb = y + (long long)b * C * T;
    // four channels per step, their loads in flight together; per channel the sum is still lane-strided partials + the same
    // shuffle tree (bit-identical to one channel at a time)
    for (int c0 = warp; c0 < C; c0 += 4 * nwarp) {
        float sq[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = lane; t < T; t += 32) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(yb + (long long)min(c0 + u * nwarp, C - 1) * T + t);
#pragma unroll
            for (int u = 0; u < 4; ++u) sq[u] = fmaf(v[u], v[u], sq[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float s = sq[u];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && c0 + u * nwarp < C) g[c0 + u * nwarp] = sqrtf(s);
        }
    }
    __syncthreads();
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += g[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) part[warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < nwarp; ++i) tot += part[i];
    const float denom = tot / (float)C + 1e-6f;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        scale[(long long)b * C + c] = fmaf(__ldg(gamma + c), g[c] / denom, 1.0f);
}

int grn_scale(const float* y, const float* gamma, float* scale, int B, int C, int T, cudaStream_t s) {
    grn_scale_kernel<<<B, 256, sizeof(float) * (C + 8), s>>>(y, gamma, scale, C, T);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// pitch_decode (encoder.py:48-67): top-4 of the 512 logits per frame, softmax over those four,
// f0 = sum p_i * 20*2^(id_i/48) with frequencies <= 20 Hz zeroed, and the result zeroed if <= 20.
// ---------------------------------------------------------------------------------------------
// One warp per frame: lane l scans classes l, l + 32, ... (all loads in flight at once) keeping its own sorted top-4, then
// four rounds of a warp arg-max (value descending, class index ascending on ties) pop the global top-4 in the order the
// sequential scan of one thread would have produced; lane 0 evaluates the softmax-weighted frequency.
__global__ void __launch_bounds__(256) pitch_decode_kernel(const float* __restrict__ logits, float* __restrict__ f0, int ncls, int T,
                                                           long long ncol) {
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= ncol) return;
    const long long b = n / T;
    const int t = (int)(n - b * T);
    const float* lp = logits + b * ncls * (long long)T + t;
    float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int id[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int c0 = 0; c0 < ncls; c0 += 32 * 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 32 + lane;
            x[u] = c < ncls ? __ldg(lp + (long long)c * T) : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 32 + lane;
            if (x[u] > v[3]) {                       // classes arrive in increasing order: strict '>' keeps the lower index
                v[3] = x[u]; id[3] = c;
#pragma unroll
                for (int q = 3; q > 0; --q)
                    if (v[q] > v[q - 1]) {
                        const float tv = v[q]; v[q] = v[q - 1]; v[q - 1] = tv;
                        const int ti = id[q]; id[q] = id[q - 1]; id[q - 1] = ti;
                    }
            }
        }
    }
    float tv[4];
    int ti[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float bv = v[0];
        int bi = id[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        tv[r] = bv;
        ti[r] = bi == 0x7fffffff ? 0 : bi;           // fewer than four comparable logits (NaN / -inf rows): class 0 at -inf, as before
        if (id[0] == bi && bi != 0x7fffffff) {       // the owner pops its head
            v[0] = v[1]; id[0] = id[1]; v[1] = v[2]; id[1] = id[2]; v[2] = v[3]; id[2] = id[3];
            v[3] = -INFINITY; id[3] = 0x7fffffff;
        }
    }
    if (lane != 0) return;
    float e[4], se = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        e[i] = expf(tv[i] - tv[0]);
        se += e[i];
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float fr = 20.0f * exp2f((float)ti[i] / 48.0f);
        if (fr <= 20.0f) fr = 0.f;
        acc = __fadd_rn(acc, __fmul_rn(e[i] / se, fr));
    }
    f0[n] = acc <= 20.0f ? 0.f : acc;
}

int pitch_decode(const float* logits, float* f0, int B, int ncls, int T, cudaStream_t s) {
    const long long ncol = (long long)B * T;
    pitch_decode_kernel<<<(unsigned)cdiv(ncol * 32, 256), 256, 0, s>>>(logits, f0, ncls, T, ncol);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// interp_linear: y[r, u] = F.interpolate(x, mode='linear')[r, u] for R independent rows
// (decoder.py:148,174;  exact coordinate arithmetic in tvc_common.cuh).
// ---------------------------------------------------------------------------------------------
__global__ void interp_linear_kernel(const float* __restrict__ x, float* __restrict__ y, int tin, int tout, float scale,
                                     long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long r = i / tout;
    const int u = (int)(i - r * tout);
    const LinCoord c = lin_coord(u, scale, tin);
    const float* xr = x + r * tin;
    y[i] = lin_blend(__ldg(xr + c.i0), __ldg(xr + c.i1), c);
}

int interp_linear(const float* x, float* y, long long rows, int tin, int tout, float scale, cudaStream_t s) {
    const long long total = rows * tout;
    interp_linear_kernel<<<cdiv(total, 256), 256, 0, s>>>(x, y, tin, tout, scale, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// out_conv_k7: FilterNet.output_layer, Conv1d(24 -> 1, k=7, replicate pad 3)  (decoder.py:220,233)
// Block = 256 consecutive samples of one utterance; the [C][256+6] input tile is staged in smem.
// ---------------------------------------------------------------------------------------------
constexpr int kOutTile = 256;
__global__ void __launch_bounds__(kOutTile) out_conv_k7_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ y,
                                                               int C, int T) {
    extern __shared__ float sm[];
    float* tile = sm;                          // [C][kOutTile+6]
    float* ws = sm + (size_t)C * (kOutTile + 6);   // [C][7]
    const int b = blockIdx.y, t0 = blockIdx.x * kOutTile;
    const float* xb = x + (long long)b * C * T;
    for (int e = threadIdx.x; e < C * (kOutTile + 6); e += blockDim.x) {
        const int c = e / (kOutTile + 6), j = e - c * (kOutTile + 6);
        int t = t0 + j - 3;
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        tile[e] = __ldg(xb + (long long)c * T + t);
    }
    for (int e = threadIdx.x; e < C * 7; e += blockDim.x) ws[e] = __ldg(w + e);
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= T) return;
    float acc = __ldg(bias);
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int j = 0; j < 7; ++j) acc = fmaf(ws[c * 7 + j], tile[c * (kOutTile + 6) + threadIdx.x + j], acc);
    }
    y[(long long)b * T + t] = acc;
}

int out_conv_k7(const float* x, const float* w, const float* bias, float* y, int B, int C, int T, cudaStream_t s) {
    dim3 grid(cdiv(T, kOutTile), B);
    const size_t smem = sizeof(float) * ((size_t)C * (kOutTile + 6) + C * 7);
    out_conv_k7_kernel<<<grid, kOutTile, smem, s>>>(x, w, bias, y, C, T);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// repack_conv_weight: torch [Cout][Cin][K]  ->  packed [K][Cin][CoutP] at channel offset co_off
// (one-time, at weight load).
// ---------------------------------------------------------------------------------------------
__global__ void repack_conv_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int K,
                                   int CoutP, int co_off) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)Cout * Cin * K;
    if (i >= total) return;
    const int k = (int)(i % K);
    const long long r = i / K;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    dst[((long long)k * Cin + ci) * CoutP + co_off + co] = src[i];
}

int repack_conv_weight(const float* src, float* dst, int Cout, int Cin, int K, int CoutP, int co_off, cudaStream_t s) {
    const long long total = (long long)Cout * Cin * K;
    repack_conv_kernel<<<cdiv(total, 256), 256, 0, s>>>(src, dst, Cout, Cin, K, CoutP, co_off);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
