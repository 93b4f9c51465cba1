// STFT magnitude by a shared-memory FFT (utils/spectrogram.py:8-15: torch.stft(n_fft = 1920, hop = 480, hann, centred,
// reflect padding).abs()[:, :, 1:]).
//
// 1920 = 15 x 16 x 8.  Two real frames ride one complex transform (z = a + i b; A[k] = (Z[k] + conj Z[N-k]) / 2,
// B[k] = (Z[k] - conj Z[N-k]) / 2i), evaluated as three decimation stages whose small DFTs (15, 16, 8 points) are done
// whole by one thread from registers with their coefficients as instruction immediates:
//   stage 1  n = 128 n1 + n2:      Y[k1][n2]  = W_1920^(n2 k1)  * sum_n1 x[128 n1 + n2] W_15^(n1 k1)        128 DFT-15
//   stage 2  n2 = 8 m1 + m2:       U[k1][j1][m2] = W_128^(m2 j1) * sum_m1 Y[k1][8 m1 + m2] W_16^(m1 j1)      120 DFT-16
//   stage 3  k = k1 + 15 j1 + 240 j2:  X[k] = sum_m2 U[k1][j1][m2] W_8^(m2 j2)                                240 DFT-8
// ~150 k FMA per frame instead of the 7.4 M of the DFT-as-GEMM it replaces.  A CTA of 256 threads owns 8 consecutive frames
// of one utterance: two transforms at a time (128 threads each), twice; the magnitudes wait in shared memory so that the
// channels-first output [B][961][Lf] is written as 32-byte runs.
#include "nets.cuh"

namespace tvc {

namespace {

constexpr int kN = kNfft;             // 1920
constexpr int kS = 136;               // padded row stride (complex elements) of the [15][128] intermediate layouts
constexpr int kFramesPerCta = 8;
struct C2 { float x, y; };
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ float2 cfma(float2 a, C2 w, float2 acc) {      // acc + a * w
    acc.x = fmaf(a.x, w.x, acc.x); acc.x = fmaf(-a.y, w.y, acc.x);
    acc.y = fmaf(a.x, w.y, acc.y); acc.y = fmaf(a.y, w.x, acc.y);
    return acc;
}
// e^(-2 pi i j / r), fp64 values rounded once
__device__ constexpr C2 kW15[15] = {{1.000000000e+00f, -0.000000000e+00f}, {9.135454576e-01f, -4.067366431e-01f}, {6.691306064e-01f, -7.431448255e-01f}, {3.090169944e-01f, -9.510565163e-01f}, {-1.045284633e-01f, -9.945218954e-01f}, {-5.000000000e-01f, -8.660254038e-01f}, {-8.090169944e-01f, -5.877852523e-01f}, {-9.781476007e-01f, -2.079116908e-01f}, {-9.781476007e-01f, 2.079116908e-01f}, {-8.090169944e-01f, 5.877852523e-01f}, {-5.000000000e-01f, 8.660254038e-01f}, {-1.045284633e-01f, 9.945218954e-01f}, {3.090169944e-01f, 9.510565163e-01f}, {6.691306064e-01f, 7.431448255e-01f}, {9.135454576e-01f, 4.067366431e-01f}};
__device__ constexpr C2 kW16[16] = {{1.000000000e+00f, -0.000000000e+00f}, {9.238795325e-01f, -3.826834324e-01f}, {7.071067812e-01f, -7.071067812e-01f}, {3.826834324e-01f, -9.238795325e-01f}, {6.123233996e-17f, -1.000000000e+00f}, {-3.826834324e-01f, -9.238795325e-01f}, {-7.071067812e-01f, -7.071067812e-01f}, {-9.238795325e-01f, -3.826834324e-01f}, {-1.000000000e+00f, -1.224646799e-16f}, {-9.238795325e-01f, 3.826834324e-01f}, {-7.071067812e-01f, 7.071067812e-01f}, {-3.826834324e-01f, 9.238795325e-01f}, {-1.836970199e-16f, 1.000000000e+00f}, {3.826834324e-01f, 9.238795325e-01f}, {7.071067812e-01f, 7.071067812e-01f}, {9.238795325e-01f, 3.826834324e-01f}};
__device__ constexpr C2 kW8[8] = {{1.000000000e+00f, -0.000000000e+00f}, {7.071067812e-01f, -7.071067812e-01f}, {6.123233996e-17f, -1.000000000e+00f}, {-7.071067812e-01f, -7.071067812e-01f}, {-1.000000000e+00f, -1.224646799e-16f}, {-7.071067812e-01f, 7.071067812e-01f}, {-1.836970199e-16f, 1.000000000e+00f}, {7.071067812e-01f, 7.071067812e-01f}};

// one small DFT: out[k] = sum_j in[j] W_r^(j k), coefficients folded into the instructions by full unrolling
template <int R>
__device__ __forceinline__ C2 wr(int i) {
    if constexpr (R == 15) return kW15[i];
    else if constexpr (R == 16) return kW16[i];
    else return kW8[i];
}
template <int R>
__device__ __forceinline__ float2 dft_out(const float2 (&in)[R], int k) {
    float2 acc = in[0];
#pragma unroll
    for (int j = 1; j < R; ++j) acc = cfma(in[j], wr<R>((j * k) % R), acc);
    return acc;
}

__global__ void __launch_bounds__(256) stft_fft_kernel(const float* __restrict__ wf, const float* __restrict__ window,
                                                       const float2* __restrict__ tw,      // W_1920^m, m < 1920
                                                       float* __restrict__ spec, int L, int Lf, int groups_per_utt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tws = reinterpret_cast<float2*>(smem_raw);                       // [1920]
    float2* bufA = tws + kN;                                                  // [2 transforms][15 * kS]
    float2* bufB = bufA + 2 * 15 * kS;                                        // [2 transforms][15 * kS]
    float* mag = reinterpret_cast<float*>(bufB + 2 * 15 * kS);                // [961][8]
    const int tid = threadIdx.x, half = tid >> 7, lt = tid & 127;
    const int b = blockIdx.x / groups_per_utt, t0 = (blockIdx.x - b * groups_per_utt) * kFramesPerCta;
    const float* x = wf + (long long)b * L;
    for (int i = tid; i < kN; i += 256) tws[i] = __ldg(tw + i);
    float2* A = bufA + half * 15 * kS;
    float2* Bf = bufB + half * 15 * kS;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int fa = t0 + pass * 4 + half * 2, fb = fa + 1;                 // the two frames of this transform
        __syncthreads();                                                      // buffers free (and the twiddles in place)
        // ---- load: z[n] = a[n] + i b[n], windowed, reflect-padded; laid out [n1][n2] with row stride kS
        for (int n = lt; n < kN; n += 128) {
            const float w = __ldg(window + n);
            float2 z = make_float2(0.f, 0.f);
            if (fa < Lf) {
                int j = kFrame * fa - kFrame + n;
                j = j < 0 ? -j : (j >= L ? 2 * (L - 1) - j : j);
                z.x = __fmul_rn(__ldg(x + j), w);
            }
            if (fb < Lf) {
                int j = kFrame * fb - kFrame + n;
                j = j < 0 ? -j : (j >= L ? 2 * (L - 1) - j : j);
                z.y = __fmul_rn(__ldg(x + j), w);
            }
            A[(n >> 7) * kS + (n & 127)] = z;
        }
        __syncthreads();
        // ---- stage 1: thread n2 = lt; DFT-15 over n1, twiddle W_1920^(n2 k1) -> B[k1][n2]
        {
            float2 in[15];
#pragma unroll
            for (int n1 = 0; n1 < 15; ++n1) in[n1] = A[n1 * kS + lt];
#pragma unroll
            for (int k1 = 0; k1 < 15; ++k1) {
                const float2 y = dft_out<15>(in, k1);
                Bf[k1 * kS + lt] = cmul(y, tws[(lt * k1) % kN]);
            }
        }
        __syncthreads();
        // ---- stage 2: thread (k1, m2), 120 of them; DFT-16 over m1, twiddle W_128^(m2 j1) -> A[k1][m2 * 17 + j1]
        if (lt < 120) {
            const int k1 = lt >> 3, m2 = lt & 7;
            float2 in[16];
#pragma unroll
            for (int m1 = 0; m1 < 16; ++m1) in[m1] = Bf[k1 * kS + 8 * m1 + m2];
#pragma unroll
            for (int j1 = 0; j1 < 16; ++j1) {
                const float2 u = dft_out<16>(in, j1);
                A[k1 * kS + m2 * 17 + j1] = cmul(u, tws[15 * m2 * j1]);
            }
        }
        __syncthreads();
        // ---- stage 3: thread (k1, j1), 240 of them (two per thread); DFT-8 over m2 -> X[k1 + 15 j1 + 240 j2] in B (flat)
        for (int q = lt; q < 240; q += 128) {
            const int k1 = q >> 4, j1 = q & 15;
            float2 in[8];
#pragma unroll
            for (int m2 = 0; m2 < 8; ++m2) in[m2] = A[k1 * kS + m2 * 17 + j1];
#pragma unroll
            for (int j2 = 0; j2 < 8; ++j2) Bf[k1 + 15 * j1 + 240 * j2] = dft_out<8>(in, j2);
        }
        __syncthreads();
        // ---- split the two real spectra, magnitudes of bins 0 .. 960 -> mag[k][frame slot]
        for (int k = lt; k < kBins; k += 128) {
            const float2 zk = Bf[k], zn = Bf[(kN - k) % kN];
            const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);       // A = (Z[k] + conj Z[N-k]) / 2
            const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);      // B = (Z[k] - conj Z[N-k]) / 2i
            const int slot = pass * 4 + half * 2;
            mag[k * kFramesPerCta + slot] = hypotf(ar, ai);
            mag[k * kFramesPerCta + slot + 1] = hypotf(br, bi);
        }
    }
    __syncthreads();
    // ---- channels-first store: spec[b][k][t0 .. t0 + 8)
    const int nf = Lf - t0 < kFramesPerCta ? Lf - t0 : kFramesPerCta;
    for (int i = tid; i < kBins * kFramesPerCta; i += 256) {
        const int k = i >> 3, f = i & 7;
        if (f < nf) spec[((long long)b * kBins + k) * Lf + t0 + f] = mag[i];
    }
}

constexpr size_t kFftSmem = sizeof(float2) * (kN + 4 * 15 * kS) + sizeof(float) * kBins * kFramesPerCta;

__global__ void stft_twiddle_kernel(float2* tw) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= kN) return;
    double s, c;
    sincospi(-2.0 * (double)m / (double)kN, &s, &c);
    tw[m] = make_float2((float)c, (float)s);
}

}  // namespace

int stft_fft_init(float2** tw_out) {
    float2* tw = nullptr;
    TVC_CUDA(cudaMalloc(&tw, sizeof(float2) * kN));
    stft_twiddle_kernel<<<cdiv(kN, 256), 256>>>(tw);
    TVC_LAUNCH_CHECK();
    TVC_CUDA(cudaFuncSetAttribute(stft_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFftSmem));
    TVC_CUDA(cudaDeviceSynchronize());
    *tw_out = tw;
    return 0;
}

int stft_fft_launch(const float* wf, const float* window, const float2* tw, float* spec, int B, int L, int Lf, cudaStream_t s) {
    const int groups = cdiv(Lf, kFramesPerCta);
    stft_fft_kernel<<<(unsigned)((long long)B * groups), 256, kFftSmem, s>>>(wf, window, tw, spec, L, Lf, groups);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
