// Signal front end of Generator.convert (module/utils): energy envelope, STFT magnitude, pitch shift.
#include <mutex>
#include <vector>
#include <cmath>

#include "nets.cuh"

namespace tvc {

// ---------------------------------------------------------------------------------------------
// estimate_energy (utils/energy_estimation.py:9-14):
//   pooled[b,i] = max_{j in [64i-32, 64i+96) & [0,L)} |wf[b,j]|      F.max_pool1d(|x|, 128, 64, 32)
//   energy      = F.interpolate(pooled, L, mode='linear')             (interp_linear, scale = P/L)
// one warp per pooling window.
// ---------------------------------------------------------------------------------------------
__global__ void energy_pool_kernel(const float* __restrict__ wf, float* __restrict__ pooled, int L, int P,
                                   long long nwin) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nwin) return;
    const long long b = w / P;
    const int i = (int)(w - b * P);
    const float* x = wf + b * L;
    float m = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = 64 * i - 32 + q * 32 + lane;
        if (j >= 0 && j < L) m = fmaxf(m, fabsf(__ldg(x + j)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) pooled[w] = m;
}

int energy_pooled_len(int L) { return (L + 64 - 128) / 64 + 1; }

int energy_estimate(const float* wf, float* energy, float* pooled, int B, int L, cudaStream_t s) {
    TVC_REQUIRE(L >= 64, "estimate_energy: need at least 64 samples, got %d", L);
    const int P = energy_pooled_len(L);
    const long long nwin = (long long)B * P;
    energy_pool_kernel<<<cdiv(nwin * 32, 256), 256, 0, s>>>(wf, pooled, L, P, nwin);
    TVC_LAUNCH_CHECK();
    const float scale = (float)P / (float)L;
    return interp_linear(pooled, energy, B, P, L, scale, s);
}

// ---------------------------------------------------------------------------------------------
// spectrogram (utils/spectrogram.py:8-15): torch.stft(n_fft=1920, hop=480, hann, center=True,
// pad_mode='reflect').abs()[:, :, 1:].  Frame t = 1..Lf covers samples 480t-960+k, k<1920, of
// the reflect-padded signal.  One kernel: windowing, a 1920-point shared-memory FFT (two real frames per complex
// transform) and the magnitudes, written channels-first (stft_fft.cu).
// ---------------------------------------------------------------------------------------------
int stft_fft_init(float2** tw_out);
int stft_fft_launch(const float* wf, const float* window, const float2* tw, float* spec, int B, int L, int Lf, cudaStream_t s);

struct StftBasis {
    float* window = nullptr;   // [1920] periodic Hann
    float2* tw = nullptr;      // [1920] e^(-2 pi i m / 1920)
};
static std::mutex g_stft_mu;
static StftBasis g_stft[64];
static bool g_stft_ready[64] = {false};

static int stft_basis(const StftBasis** out) {
    int dev = 0;
    TVC_CUDA(cudaGetDevice(&dev));
    TVC_REQUIRE(dev >= 0 && dev < 64, "spectrogram: unsupported device ordinal %d", dev);
    std::lock_guard<std::mutex> lock(g_stft_mu);
    if (!g_stft_ready[dev]) {
        std::vector<float> win(kNfft);
        for (int k = 0; k < kNfft; ++k) win[k] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)k / (double)kNfft));   // periodic Hann
        float* dwin = nullptr;
        TVC_CUDA(cudaMalloc(&dwin, sizeof(float) * win.size()));
        TVC_CUDA(cudaMemcpy(dwin, win.data(), sizeof(float) * win.size(), cudaMemcpyHostToDevice));
        g_stft[dev].window = dwin;
        TVC_TRY(stft_fft_init(&g_stft[dev].tw));
        g_stft_ready[dev] = true;
    }
    *out = &g_stft[dev];
    return 0;
}

int spectrogram_run(Arena& A, cudaStream_t s, const float* wf, float* spec, int B, int L) {
    TVC_REQUIRE(L % kFrame == 0, "spectrogram: L=%d is not a multiple of %d (autopad first)", L, kFrame);
    TVC_REQUIRE(L > kNfft / 2, "spectrogram: reflect padding needs more than %d samples, got %d", kNfft / 2, L);
    if (A.dry) return 0;
    const StftBasis* sb = nullptr;
    TVC_TRY(stft_basis(&sb));
    ProfScope ps("stft_fft(", s);
    return stft_fft_launch(wf, sb->window, sb->tw, spec, B, L, L / kFrame, s);
}

// ---------------------------------------------------------------------------------------------
// shift_frequency (utils/pitch_shift.py:5-15), op for op in fp32:
//   midi = log2(relu(f/440) + 1e-6) * 12 + 69 ; midi += shift ; out = 440 * 2^((midi - 69)/12)
// ---------------------------------------------------------------------------------------------
__global__ void shift_frequency_kernel(const float* __restrict__ f0, float* __restrict__ out, long long n, float shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = __fdiv_rn(f0[i], 440.0f);
    const float r = __fadd_rn(fmaxf(a, 0.f), 1e-6f);
    float m = __fadd_rn(__fmul_rn(log2f(r), 12.0f), 69.0f);
    m = __fadd_rn(m, shift);
    const float e = __fdiv_rn(__fsub_rn(m, 69.0f), 12.0f);
    out[i] = __fmul_rn(440.0f, exp2f(e));
}

int shift_frequency(const float* f0, float* out, long long n, float shift, cudaStream_t s) {
    if (n <= 0) return 0;
    shift_frequency_kernel<<<cdiv(n, 256), 256, 0, s>>>(f0, out, n, shift);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
