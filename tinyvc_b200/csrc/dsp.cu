// Harmonic-plus-noise source (Decoder.dsp, decoder.py:24-85,259-266) as sm_100a kernels.
#include "tvc_kernels.cuh"

namespace tvc {

// ---------------------------------------------------------------------------------------------
// Harmonic oscillator.  For utterance b, oscillator k = 1..15, sample n:
//     fs   = interp(f0[b], L)[n] * k                 fp32          (decoder.py:39,42)
//     inc  = fs / 24000                              fp32 true division (:49)
//     I    = fp32( sum_{m<=n} fp64(inc[m]) )         torch.cumsum on CPU accumulates fp32 inputs
//                                                    in fp64, sequentially, rounding each prefix
//     th   = fp32(2*pi) * fmodf(I, 1)                (:50)
//     h    = sinf(th) * interp(f0 > 20)[n]           (:45-46,52)
//     src  = h * interp(amps[b,k-1], x480)[n]        (:262-263)
// The phase I reaches 1e4..1e6 cycles, where one fp32 ulp is a visible fraction of a cycle, so
// the rounding sequence above -- including the *sequential* order of the fp64 sum -- is part of
// the contract.  One warp owns one (b,k) row: the 32 lanes compute 32 consecutive increments in
// parallel (interp, multiply, IEEE divide), then every lane replays the same 32-step fp64 chain
// over warp shuffles and keeps the prefix belonging to its own sample.  The chain (one DADD per
// sample) is the only serial part; everything else, and the store, is coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOsc * 32) harmonic_osc_kernel(const float* __restrict__ f0,     // [B][Lf]
                                                                 const float* __restrict__ amps,   // [B][15][Lf]
                                                                 float* __restrict__ src, long long src_bs,
                                                                 int Lf, float scale_size, float scale_factor) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int k = (threadIdx.x >> 5) + 1;
    const int L = Lf * kFrame;
    const float* f0b = f0 + (long long)b * Lf;
    const float* ab = amps + ((long long)b * kOsc + (k - 1)) * Lf;
    float* out = src + (long long)b * src_bs + (long long)(k - 1) * L;
    const float kf = (float)k;
    double acc = 0.0;
    for (int base = 0; base < L; base += 32) {
        const int n = base + lane;
        const LinCoord c = lin_coord(n, scale_size, Lf);
        const float fa = __ldg(f0b + c.i0), fb = __ldg(f0b + c.i1);
        const float fs = __fmul_rn(lin_blend(fa, fb, c), kf);
        const double incd = (double)__fdiv_rn(fs, kSampleRate);
        double mine = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, incd, j));
            if (j == lane) mine = acc;
        }
        const float I = __double2float_rn(mine);
        const float theta = __fmul_rn(6.28318530717958647692f, fmodf(I, 1.0f));
        const float uv = lin_blend(fa > 20.0f ? 1.f : 0.f, fb > 20.0f ? 1.f : 0.f, c);
        const float h = __fmul_rn(sinf(theta), uv);
        const LinCoord ca = lin_coord(n, scale_factor, Lf);
        const float a = lin_blend(__ldg(ab + ca.i0), __ldg(ab + ca.i1), ca);
        out[n] = __fmul_rn(h, a);
    }
}

int harmonic_osc(const float* f0, const float* amps, float* src, long long src_bs, int B, int Lf, cudaStream_t s) {
    const int L = Lf * kFrame;
    const float scale_size = (float)Lf / (float)L;              // F.interpolate(size=L): float(in)/float(out)
    const float scale_factor = (float)(1.0 / (double)kFrame);   // F.interpolate(scale_factor=480): float(1/sf)
    harmonic_osc_kernel<<<B, kOsc * 32, 0, s>>>(f0, amps, src, src_bs, Lf, scale_size, scale_factor);
    TVC_LAUNCH_CHECK();
    return 0;
}

// Debug/parity helper: theta only (no uv, no amps) for one oscillator row layout [B][15][L].
__global__ void __launch_bounds__(kOsc * 32) harmonic_theta_kernel(const float* __restrict__ f0, float* __restrict__ theta,
                                                                   int Lf, float scale_size) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int k = (threadIdx.x >> 5) + 1;
    const int L = Lf * kFrame;
    const float* f0b = f0 + (long long)b * Lf;
    float* out = theta + ((long long)b * kOsc + (k - 1)) * L;
    double acc = 0.0;
    for (int base = 0; base < L; base += 32) {
        const int n = base + lane;
        const LinCoord c = lin_coord(n, scale_size, Lf);
        const float fs = __fmul_rn(lin_blend(__ldg(f0b + c.i0), __ldg(f0b + c.i1), c), (float)k);
        const double incd = (double)__fdiv_rn(fs, kSampleRate);
        double mine = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, incd, j));
            if (j == lane) mine = acc;
        }
        out[n] = __fmul_rn(6.28318530717958647692f, fmodf(__double2float_rn(mine), 1.0f));
    }
}

int harmonic_theta(const float* f0, float* theta, int B, int Lf, cudaStream_t s) {
    const float scale_size = (float)Lf / (float)(Lf * kFrame);
    harmonic_theta_kernel<<<B, kOsc * 32, 0, s>>>(f0, theta, Lf, scale_size);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Noise branch (decoder.py:63-85).
//   angle = (rand01*2)*pi - pi ;  Y = kernel * exp(j*angle)                         (:78-80)
//   frames = irfft(Y, 1920) for Lf frames preceded by one all-zero frame            (:81-82)
//   noise  = overlap-add(hop 480, rectangular window) / coverage, centre-trimmed    (torch.istft)
// The inverse real DFT is evaluated as two dense products against a precomputed basis using
// x[p] = C[p] - S[p], x[N-p] = C[p] + S[p]  (C = cos-part of Re, S = sin-part of Im, p <= N/2),
// which halves the work of a plain 1920x1922 product and runs on the generic conv kernel.
// ---------------------------------------------------------------------------------------------
__global__ void noise_spectrum_kernel(const float* __restrict__ kern, const float* __restrict__ rand01,
                                      float* __restrict__ yri, int Lf, long long per_b /* 961*Lf */, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / per_b, r = i - b * per_b;
    const float pi_f = 3.14159265358979323846f;
    const float a = __fsub_rn(__fmul_rn(__fmul_rn(__ldg(rand01 + i), 2.0f), pi_f), pi_f);
    float sn, cs;
    sincosf(a, &sn, &cs);
    const float kv = __ldg(kern + i);
    float* yb = yri + b * 2 * per_b;          // [B][2*961][Lf]: rows [0,961) real, [961,1922) imag
    yb[r] = __fmul_rn(cs, kv);
    yb[per_b + r] = __fmul_rn(sn, kv);
}

int noise_spectrum(const float* kern, const float* rand01, float* yri, int B, int Lf, cudaStream_t s) {
    const long long per_b = (long long)kBins * Lf, total = per_b * B;
    noise_spectrum_kernel<<<cdiv(total, 256), 256, 0, s>>>(kern, rand01, yri, Lf, per_b, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

// cs: [B][2*961][Lf]  rows [0,961) = C[p], rows [961,1922) = S[p].  Writes channel `ch` of src.
__global__ void noise_ola_kernel(const float* __restrict__ cs, float* __restrict__ src, long long src_bs, int ch, int Lf,
                                 long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int L = Lf * kFrame;
    const long long b = i / L;
    const int n = (int)(i - b * L);
    const int j = n / kFrame, r = n - j * kFrame;
    const float* cb = cs + b * 2 * kBins * (long long)Lf;
    const float* sb = cb + (long long)kBins * Lf;
    // frames t (0 = the prepended zero frame) covering output position n + 960
    const int tlo = j - 1 < 0 ? 0 : j - 1;
    const int thi = j + 2 > Lf ? Lf : j + 2;
    float acc = 0.f;
    for (int t = tlo; t <= thi; ++t) {
        if (t == 0) continue;                      // zero frame: contributes 0 but counts in the envelope
        const int q = kFrame * (j + 2 - t) + r;    // position inside frame t, 0..1919
        const int col = t - 1;
        float v;
        if (q <= kNfft / 2) v = __fsub_rn(__ldg(cb + (long long)q * Lf + col), __ldg(sb + (long long)q * Lf + col));
        else v = __fadd_rn(__ldg(cb + (long long)(kNfft - q) * Lf + col), __ldg(sb + (long long)(kNfft - q) * Lf + col));
        acc = __fadd_rn(acc, v);
    }
    src[b * src_bs + (long long)ch * L + n] = __fdiv_rn(acc, (float)(thi - tlo + 1));
}

int noise_ola(const float* cs, float* src, long long src_bs, int ch, int B, int Lf, cudaStream_t s) {
    const long long total = (long long)B * Lf * kFrame;
    noise_ola_kernel<<<cdiv(total, 256), 256, 0, s>>>(cs, src, src_bs, ch, Lf, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
