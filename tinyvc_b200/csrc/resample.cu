// Sample-rate conversion of input files (reference infer.py:45-46,63-64: torchaudio.functional.resample(wf, sr, 24000) with
// its defaults -- lowpass_filter_width 6, rolloff 0.99, "sinc_interp_hann").  torchaudio is a third-party dependency of the
// reference (requirements.txt, un-pinned; 2.11.0 in this image); its published algorithm is restated here:
//
//   g = gcd(orig, new);  o = orig / g;  n = new / g;  base = min(o, n) * rolloff;  width = ceil(6 * o / base)
//   kernel[p][k] = sinc(pi * t) * cos^2(pi * t / 12) * base / o,   t = clamp((-p / n + (k - width) / o) * base, -6, 6)
//                  for p < n, k < 2 * width + o           (evaluated in fp64, rounded to fp32, like _get_sinc_resample_kernel)
//   y[j * n + p] = sum_k xpad[j * o + k] * kernel[p][k],   xpad = x zero-padded by `width` in front     (conv1d, stride o)
//   len(y)       = ceil(n * L / o)
//
// The polyphase bank (n x (2 width + o) floats: 80 x 171 for 44.1 kHz -> 24 kHz) is built on the host once per
// (device, orig, new) and kept on the device; the kernel is one thread per output sample, the bank phase-major so that
// the threads of a warp (consecutive outputs = consecutive phases) stream their own rows and share the input window in L1.
#include <cmath>
#include <map>
#include <mutex>
#include <numeric>
#include <tuple>
#include <vector>

#include "tvc_kernels.cuh"

namespace tvc {

namespace {

struct Bank {
    float* w = nullptr;   // [n][klen]
    int o = 0, n = 0, width = 0, klen = 0;
};
std::mutex g_bank_mu;
std::map<std::tuple<int, int, int>, Bank> g_banks;   // (device, orig, new)

int get_bank(int orig, int neu, Bank* out) {
    int dev = 0;
    TVC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_bank_mu);
    const auto key = std::make_tuple(dev, orig, neu);
    auto it = g_banks.find(key);
    if (it != g_banks.end()) { *out = it->second; return 0; }
    const int g = std::gcd(orig, neu);
    Bank b;
    b.o = orig / g; b.n = neu / g;
    const double rolloff = 0.99, lpw = 6.0;
    const double base = (double)std::min(b.o, b.n) * rolloff;
    b.width = (int)std::ceil(lpw * b.o / base);
    b.klen = 2 * b.width + b.o;
    TVC_REQUIRE((long long)b.n * b.klen <= (1 << 24), "resample: %d -> %d Hz needs a %d x %d filter bank", orig, neu, b.n, b.klen);
    std::vector<float> h((size_t)b.n * b.klen);
    const double pi = 3.141592653589793;   // math.pi
    for (int p = 0; p < b.n; ++p)
        for (int k = 0; k < b.klen; ++k) {
            // torch.arange(0, -new, -1) is an int64 tensor: its division by new_freq is carried out in fp32, the sum in fp64
            double t = ((double)((float)(-p) / (float)b.n) + (double)(k - b.width) / (double)b.o) * base;
            t = t < -lpw ? -lpw : (t > lpw ? lpw : t);
            const double c = std::cos(t * pi / lpw / 2.0);
            const double window = c * c;
            const double tp = t * pi;
            const double sinc = tp == 0.0 ? 1.0 : std::sin(tp) / tp;
            h[(size_t)p * b.klen + k] = (float)(sinc * (window * (base / (double)b.o)));
        }
    TVC_CUDA(cudaMalloc(&b.w, h.size() * sizeof(float)));
    TVC_CUDA(cudaMemcpy(b.w, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    g_banks[key] = b;
    *out = b;
    return 0;
}

__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                       long long L, long long Lout, int o, int n, int width, int klen,
                                                       long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / Lout, u = i - b * Lout;
    const long long j = u / n;
    const int p = (int)(u - j * n);
    const float* xb = x + b * L;
    const float* wp = w + (long long)p * klen;
    const long long s0 = j * o - width;              // xpad[j * o + k] = x[s0 + k]
    int k0 = s0 < 0 ? (int)(-s0) : 0;
    int k1 = s0 + klen > L ? (int)(L - s0) : klen;    // zero padding contributes nothing
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(xb + s0 + k), __ldg(wp + k), acc);
    y[i] = acc;
}

}  // namespace

long long resample_length(long long L, int orig, int neu) {
    const int g = std::gcd(orig, neu);
    const long long o = orig / g, n = neu / g;
    return (n * L + o - 1) / o;                       // ceil(new * L / orig), exact in integers
}

int resample_run(const float* x, float* y, int B, long long L, int orig, int neu, cudaStream_t s) {
    TVC_REQUIRE(orig > 0 && neu > 0, "resample: frequencies must be positive (%d -> %d)", orig, neu);
    if (orig == neu) {                                // torchaudio returns the input unchanged
        TVC_CUDA(cudaMemcpyAsync(y, x, sizeof(float) * (size_t)B * (size_t)L, cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    Bank b;
    TVC_TRY(get_bank(orig, neu, &b));
    const long long Lout = resample_length(L, orig, neu);
    const long long total = (long long)B * Lout;
    if (total == 0) return 0;
    resample_kernel<<<(unsigned)cdiv(total, 256), 256, 0, s>>>(x, b.w, y, L, Lout, b.o, b.n, b.width, b.klen, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
