// SOLA alignment + cross-fade of StreamInfer.audio_callback (module/infer/stream.py:74-95), batched
// over S independent streams and kept on the device (the reference syncs with .item() every tick).
//
// per stream:  temp = y[-(block+cross+search+delay) : -delay]
//   nom[i] = sum_{j<cross} temp[i+j] * sola[j]            F.conv1d(conv_input, sola_buffer)      (:77)
//   den[i] = sqrt(sum_{j<cross} temp[i+j]^2 + 1e-8)       F.conv1d(conv_input**2, ones)          (:78)
//   shift  = argmax_i nom[i]/den[i],  i in [0, search]                                           (:79)
//   out    = temp[shift : shift+block+cross];  out[:cross] = out[:cross]*fade_in + sola*fade_out  (:80,90-92)
//   sola'  = out[-cross:] ; return out[:block]                                                   (:94-95)
#include <cooperative_groups.h>
#include <cstdint>

#include "tvc_kernels.cuh"

namespace tvc {

__global__ void __launch_bounds__(1024) sola_kernel(const float* __restrict__ y, int y_len, float* __restrict__ sola_buf,
                                                    const float* __restrict__ fade_in, float* __restrict__ out_block,
                                                    int* __restrict__ shift_out, int block, int cross, int search,
                                                    int delay, float* __restrict__ pv_ab) {
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
    if (cross % 4 == 0 && block >= 4) {
        // four adjacent positions per thread: 3 vector loads feed 32 FMAs; per position the sums still run over j ascending
        // (the order of the scalar loop below, so the ratio and the arg-max are the same bits)
        const float4* t4 = reinterpret_cast<const float4*>(temp);
        const float4* s4 = reinterpret_cast<const float4*>(sola);
        for (int i0 = 4 * threadIdx.x; i0 <= search; i0 += 4 * blockDim.x) {
            float nom[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
            float4 a = t4[i0 >> 2];
#pragma unroll 2
            for (int j = 0; j < cross; j += 4) {
                // positions i0 + 1 .. i0 + 3 read up to 3 elements past position i0's window: at most temp[search + cross + 3],
                // inside temp[0, block + cross + search) for block >= 4
                const float4 b = t4[((i0 + j) >> 2) + 1];
                const float4 sv = s4[j >> 2];
                const float w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                const float sj[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        nom[q] = fmaf(w[q + u], sj[u], nom[q]);
                        den[q] = fmaf(w[q + u], w[q + u], den[q]);
                    }
                a = b;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q;
                if (i > search) break;
                const float r = __fdiv_rn(nom[q], sqrtf(__fadd_rn(den[q], 1e-8f)));
                if (r > bv || (r == bv && i < bi)) {
                    bv = r;
                    bi = i;
                }
            }
        }
    } else {
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
    if (pv_ab) {
        // phase-vocoder cross-fade (stream.py:83-89): hand a = old tail, b = new head to pv_* kernels, which
        // write out_block[:cross]; host side guarantees block >= cross, so the new tail is untouched audio
        for (int i = threadIdx.x; i < cross; i += blockDim.x) {
            pv_ab[((long long)s * 2) * cross + i] = sola[i];
            pv_ab[((long long)s * 2 + 1) * cross + i] = temp[sh + i];
        }
    }
    for (int i = threadIdx.x + (pv_ab ? cross : 0); i < block; i += blockDim.x) {
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

// The same tick on a cluster of four CTAs per stream (one SM each), for FEW streams: the lag search is FMA-bound on a single SM
// (1 921 lags x 1 920 taps x 2: ~30 us), so with SMs to spare the lags are dealt out in groups of four over 4 x 128 threads; every CTA leaves its best (ratio, lag) in
// CTA 0's shared memory through distributed shared memory, and CTA 0 -- after one cluster barrier -- picks the winner (same
// ordering rule, so the same lag) and does the cross-fade and the tail update.  Arithmetic per lag is the loop above.
constexpr int kSolaCluster = 4, kSolaThreads = 128;
__global__ void __cluster_dims__(kSolaCluster, 1, 1) __launch_bounds__(kSolaThreads)
    sola_cluster_kernel(const float* __restrict__ y, int y_len, float* __restrict__ sola_buf, const float* __restrict__ fade_in,
                        float* __restrict__ out_block, int* __restrict__ shift_out, int block, int cross, int search, int delay,
                        float* __restrict__ pv_ab) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float sm[];
    float* temp = sm;                              // [block+cross+search]
    float* sola = temp + block + cross + search;   // [cross]
    __shared__ float best_v[kSolaThreads / 32];
    __shared__ int best_i[kSolaThreads / 32];
    __shared__ float clu_v[kSolaCluster];          // CTA 0's copy receives every rank's candidate
    __shared__ int clu_i[kSolaCluster];
    __shared__ int s_shift;
    const int rank = (int)cluster.block_rank();
    const int s = blockIdx.x / kSolaCluster;
    const int tl = block + cross + search;
    const float* ys = y + (long long)s * y_len + (y_len - tl - delay);
    // 128 threads stage 30 KB: 16-byte loads, four in flight per thread (the loops below are latency chains otherwise)
    const float* sb = sola_buf + (long long)s * cross;
    if ((reinterpret_cast<uintptr_t>(ys) & 15) == 0 && (reinterpret_cast<uintptr_t>(sb) & 15) == 0 && tl % 4 == 0) {
#pragma unroll 4
        for (int i = threadIdx.x; i < tl / 4; i += blockDim.x) reinterpret_cast<float4*>(temp)[i] = __ldg(reinterpret_cast<const float4*>(ys) + i);
#pragma unroll 4
        for (int i = threadIdx.x; i < cross / 4; i += blockDim.x) reinterpret_cast<float4*>(sola)[i] = __ldg(reinterpret_cast<const float4*>(sb) + i);
    } else {
        for (int i = threadIdx.x; i < tl; i += blockDim.x) temp[i] = ys[i];
        for (int i = threadIdx.x; i < cross; i += blockDim.x) sola[i] = sb[i];
    }
    __syncthreads();
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    {
        const float4* t4 = reinterpret_cast<const float4*>(temp);
        const float4* s4 = reinterpret_cast<const float4*>(sola);
        for (int i0 = 4 * (rank * kSolaThreads + (int)threadIdx.x); i0 <= search; i0 += 4 * kSolaCluster * kSolaThreads) {
            float nom[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
            float4 a = t4[i0 >> 2];
#pragma unroll 2
            for (int j = 0; j < cross; j += 4) {
                const float4 b = t4[((i0 + j) >> 2) + 1];
                const float4 sv = s4[j >> 2];
                const float w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                const float sj[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        nom[q] = fmaf(w[q + u], sj[u], nom[q]);
                        den[q] = fmaf(w[q + u], w[q + u], den[q]);
                    }
                a = b;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q;
                if (i > search) break;
                const float r = __fdiv_rn(nom[q], sqrtf(__fadd_rn(den[q], 1e-8f)));
                if (r > bv || (r == bv && i < bi)) { bv = r; bi = i; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { best_v[warp] = bv; best_i[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = best_v[0];
        int ix = best_i[0];
        for (int w = 1; w < kSolaThreads / 32; ++w)
            if (best_v[w] > v || (best_v[w] == v && best_i[w] < ix)) { v = best_v[w]; ix = best_i[w]; }
        float* rv = cluster.map_shared_rank(clu_v, 0);
        int* ri = cluster.map_shared_rank(clu_i, 0);
        rv[rank] = v;
        ri[rank] = ix;
    }
    cluster.sync();                                // candidates of all four CTAs are in CTA 0's shared memory
    if (rank != 0) return;
    if (threadIdx.x == 0) {
        float v = clu_v[0];
        int ix = clu_i[0];
        for (int w = 1; w < kSolaCluster; ++w)
            if (clu_v[w] > v || (clu_v[w] == v && clu_i[w] < ix)) { v = clu_v[w]; ix = clu_i[w]; }
        if (ix == 0x7fffffff) ix = 0;              // all-NaN correlation: position 0, as in sola_kernel
        s_shift = ix;
        shift_out[s] = ix;
    }
    __syncthreads();
    const int sh = s_shift;
    if (pv_ab) {
        for (int i = threadIdx.x; i < cross; i += blockDim.x) {
            pv_ab[((long long)s * 2) * cross + i] = sola[i];
            pv_ab[((long long)s * 2 + 1) * cross + i] = temp[sh + i];
        }
    }
#pragma unroll 4
    for (int i = threadIdx.x + (pv_ab ? cross : 0); i < block; i += blockDim.x) {
        float v = temp[sh + i];
        if (i < cross) {
            const float fi = __ldg(fade_in + i);
            v = __fadd_rn(__fmul_rn(v, fi), __fmul_rn(sola[i], __fsub_rn(1.0f, fi)));
        }
        out_block[(long long)s * block + i] = v;
    }
    __syncthreads();   // every read of the old sola buffer is done
#pragma unroll 4
    for (int i = threadIdx.x; i < cross; i += blockDim.x) {
        const int p = block + i;
        float v = temp[sh + p];
        if (p < cross) {
            const float fi = __ldg(fade_in + p);
            v = __fadd_rn(__fmul_rn(v, fi), __fmul_rn(sola[p], __fsub_rn(1.0f, fi)));
        }
        sola_buf[(long long)s * cross + i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// phase_vocoder(a, b, fade_out, fade_in) (module/infer/stream.py:9-26), one CTA per stream:
//   window = sqrt(fade_out * fade_in);  fa = rfft(a*window), fb = rfft(b*window)
//   absab  = |fa| + |fb|, doubled for every bin but DC (and Nyquist when n is even)
//   dphi   = angle(fb) - angle(fa), wrapped to [-pi, pi);  w = 2*pi*f + dphi;  t = i/n
//   out[i] = a*fade_out^2 + b*fade_in^2 + sum_f absab[f] * cos(w[f]*t[i] + angle(fa)[f]) * window[i] / n
// pv_dft_kernel evaluates the two real DFTs directly (n = 1920: 961 bins x 1920 samples, fp32
// products accumulated in fp64, twiddles from a shared table) and leaves {absab, w, phia} per bin;
// pv_synth_kernel evaluates the cosine bank with the reference's fp32 rounding of the argument
// (w*t rounded, then + phia rounded: the argument reaches 6e3 rad, where that rounding is visible).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pv_dft_kernel(const float* __restrict__ ab, long long a_stride, long long b_off,
                                                      const float* __restrict__ fade_in, float* __restrict__ bins, int n) {
    extern __shared__ float sm[];
    float* xa = sm;            // [n] a * window
    float* xb = xa + n;        // [n] b * window
    float* tc = xb + n;        // [n] cos(2 pi k / n)
    float* ts = tc + n;        // [n] sin(2 pi k / n)
    const int s = blockIdx.x, nb = n / 2 + 1;
    const float* a = ab + (long long)s * a_stride;
    const float* b = a + b_off;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float fi = __ldg(fade_in + i), fo = __fsub_rn(1.0f, fi);
        const float w = sqrtf(__fmul_rn(fo, fi));
        xa[i] = __fmul_rn(a[i], w);
        xb[i] = __fmul_rn(b[i], w);
        double sv, cv;
        sincospi(2.0 * (double)i / (double)n, &sv, &cv);
        tc[i] = (float)cv;
        ts[i] = (float)sv;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < nb; f += blockDim.x) {
        double ar = 0.0, ai = 0.0, br = 0.0, bi = 0.0;
        int idx = 0;
        for (int k = 0; k < n; ++k) {
            const float c = tc[idx], sn = ts[idx];
            const float va = xa[k], vb = xb[k];
            ar += (double)(va * c); ai -= (double)(va * sn);
            br += (double)(vb * c); bi -= (double)(vb * sn);
            idx += f;
            if (idx >= n) idx -= n;
        }
        const float far = (float)ar, fai = (float)ai, fbr = (float)br, fbi = (float)bi;
        float absab = __fadd_rn(hypotf(far, fai), hypotf(fbr, fbi));
        const bool dbl = (n % 2 == 0) ? (f >= 1 && f < nb - 1) : (f >= 1);
        if (dbl) absab = __fmul_rn(absab, 2.0f);
        const float phia = atan2f(fai, far), phib = atan2f(fbi, fbr);
        float d = __fsub_rn(phib, phia);
        const float turns = floorf(__fadd_rn(__fdiv_rn(__fdiv_rn(d, 2.0f), 3.14159274f), 0.5f));
        d = __fsub_rn(d, __fmul_rn(6.2831855f, turns));
        const float w = __fadd_rn(__fmul_rn(6.2831855f, (float)f), d);
        float* o = bins + (long long)s * 3 * nb;
        o[f] = absab;
        o[nb + f] = w;
        o[2 * nb + f] = phia;
    }
}

__global__ void __launch_bounds__(256) pv_synth_kernel(const float* __restrict__ ab, long long a_stride, long long b_off,
                                                       const float* __restrict__ fade_in, const float* __restrict__ bins,
                                                       float* __restrict__ out, long long out_stride, int n) {
    extern __shared__ float sm[];
    const int s = blockIdx.y, nb = n / 2 + 1;
    const float* bs = bins + (long long)s * 3 * nb;
    for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) sm[i] = bs[i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* absab = sm;
    const float* w = sm + nb;
    const float* phia = sm + 2 * nb;
    const float t = __fdiv_rn((float)i, (float)n);
    double acc = 0.0;
    for (int f = 0; f < nb; ++f) {
        const float arg = __fadd_rn(__fmul_rn(w[f], t), phia[f]);
        acc += (double)__fmul_rn(absab[f], cosf(arg));
    }
    const float* a = ab + (long long)s * a_stride;
    const float fi = __ldg(fade_in + i), fo = __fsub_rn(1.0f, fi);
    const float win = sqrtf(__fmul_rn(fo, fi));
    const float lin = __fadd_rn(__fmul_rn(a[i], __fmul_rn(fo, fo)), __fmul_rn(a[b_off + i], __fmul_rn(fi, fi)));
    out[(long long)s * out_stride + i] = __fadd_rn(lin, __fdiv_rn(__fmul_rn((float)acc, win), (float)n));
}

size_t phase_vocoder_scratch_floats(int S, int n) { return (size_t)S * 3 * (size_t)(n / 2 + 1); }

// a = ab + s*a_stride, b = a + b_off (floats); out[s*out_stride + i], i < n; bins: phase_vocoder_scratch_floats(S, n) floats
int phase_vocoder_run(const float* ab, long long a_stride, long long b_off, const float* fade_in, float* bins, float* out,
                      long long out_stride, int S, int n, cudaStream_t s) {
    TVC_REQUIRE(n >= 2 && n <= 8192, "phase_vocoder: cross-fade length %d out of range [2, 8192]", n);
    const size_t smem = sizeof(float) * 4 * (size_t)n;
    static PerDeviceOnce attr;
    TVC_TRY(attr.run([] { TVC_CUDA(cudaFuncSetAttribute(pv_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); return 0; }));
    pv_dft_kernel<<<S, 1024, smem, s>>>(ab, a_stride, b_off, fade_in, bins, n);
    TVC_LAUNCH_CHECK();
    pv_synth_kernel<<<dim3(cdiv(n, 256), S), 256, sizeof(float) * 3 * (size_t)(n / 2 + 1), s>>>(ab, a_stride, b_off, fade_in, bins, out,
                                                                                          out_stride, n);
    TVC_LAUNCH_CHECK();
    return 0;
}

int sola_run(const float* y, int y_len, float* sola_buf, const float* fade_in, float* out_block, int* shift_out, int S,
             int block, int cross, int search, int delay, cudaStream_t s, float* pv_scratch) {
    TVC_REQUIRE(y_len >= block + cross + search + delay, "sola: window of %d samples is shorter than block+cross+search+delay = %d",
                y_len, block + cross + search + delay);
    const size_t smem = sizeof(float) * (size_t)(block + 2 * cross + search);
    TVC_REQUIRE(smem <= 200 * 1024, "sola: block/cross/search too large for shared memory (%zu bytes)", smem);
    static PerDeviceOnce attr;
    TVC_TRY(attr.run([] { TVC_CUDA(cudaFuncSetAttribute(sola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); return 0; }));
    TVC_REQUIRE(!pv_scratch || block >= cross, "sola: the phase-vocoder cross-fade needs block (%d) >= cross-fade (%d)", block, cross);
    int dev = 0, sms = 148;
    TVC_CUDA(cudaGetDevice(&dev));
    TVC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (cross % 4 == 0 && block >= 4 && S * kSolaCluster <= sms) {
        // few streams: four CTAs (one cluster) per stream (measured: one stream's tick 0.990 -> 0.956 ms; at 128 streams the
        // single-CTA kernel is faster, 55 vs 83 us)
        static PerDeviceOnce attr2;
        TVC_TRY(attr2.run([] { TVC_CUDA(cudaFuncSetAttribute(sola_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); return 0; }));
        sola_cluster_kernel<<<S * kSolaCluster, kSolaThreads, smem, s>>>(y, y_len, sola_buf, fade_in, out_block, shift_out, block, cross, search,
                                                                       delay, pv_scratch);
    } else {
        sola_kernel<<<S, 1024, smem, s>>>(y, y_len, sola_buf, fade_in, out_block, shift_out, block, cross, search, delay, pv_scratch);
    }
    TVC_LAUNCH_CHECK();
    if (pv_scratch) {
        // scratch layout: [S][2][cross] (a, b) then the per-bin table
        float* bins = pv_scratch + (size_t)S * 2 * cross;
        return phase_vocoder_run(pv_scratch, 2LL * cross, cross, fade_in, bins, out_block, block, S, cross, s);
    }
    return 0;
}

}  // namespace tvc
