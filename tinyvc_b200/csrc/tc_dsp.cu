// Harmonic-plus-noise source (Decoder.dsp, decoder.py:24-85,259-266) for the tensor-core path:
// channels-last operands, and the oscillator's fp64 cumulative phase as a three-level scan.
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

// ---------------------------------------------------------------------------------------------
// noise_spectrum_cl (decoder.py:78-80): angle = (rand01*2)*pi - pi, Y = kernel * exp(j*angle).
// rand01 arrives channels-first [B][961][Lf] (the layout torch.rand draws it in); a 32x32 tile
// transpose through shared memory turns it into channels-last rows.  Outputs: split planes of
// Re(Y) and Im(Y), [B*Lf][y_cs], channels >= 961 zero.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) noise_spectrum_cl_kernel(const float* __restrict__ kern, int k_cs,
                                                                const float* __restrict__ rand01,
                                                                bf16* __restrict__ yr_hi, bf16* __restrict__ yr_lo,
                                                                bf16* __restrict__ yi_hi, bf16* __restrict__ yi_lo,
                                                                int y_cs, int Lf) {
    TVC_PDL_PROLOGUE();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, t = t0 + tx;
        tile[j][tx] = (c < kBins && t < Lf) ? __ldg(rand01 + ((long long)b * kBins + c) * Lf + t) : 0.f;
    }
    __syncthreads();
    const float pi_f = 3.14159265358979323846f;
    for (int j = ty; j < 32; j += 8) {
        const int t = t0 + j, c = c0 + tx;
        if (t >= Lf || c >= y_cs) continue;
        const long long row = (long long)b * Lf + t;
        float re = 0.f, im = 0.f;
        if (c < kBins) {
            const float a = __fsub_rn(__fmul_rn(__fmul_rn(tile[tx][j], 2.0f), pi_f), pi_f);
            float sn, cs;
            sincosf(a, &sn, &cs);
            const float kv = __ldg(kern + row * k_cs + c);
            re = __fmul_rn(cs, kv);
            im = __fmul_rn(sn, kv);
        }
        bf16 h, l;
        split_bf16(re, h, l);
        yr_hi[row * y_cs + c] = h; yr_lo[row * y_cs + c] = l;
        split_bf16(im, h, l);
        yi_hi[row * y_cs + c] = h; yi_lo[row * y_cs + c] = l;
    }
}

// ---------------------------------------------------------------------------------------------
// noise_ola_cl (decoder.py:81-82 = torch.istft, n_fft 1920, hop 480, rectangular window, centred,
// one all-zero frame prepended).  The inverse real DFT of frame t is x[p] = C[p] - S[p],
// x[N-p] = C[p] + S[p] (p <= N/2) with C = cos-basis * Re(Y), S = sin-basis * Im(Y) computed by
// two tensor-core products; this kernel overlap-adds and divides by the window coverage.
// c, sn: fp32 [B*Lf][cs];  noise: [B][L].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) noise_ola_cl_kernel(const float* __restrict__ c, const float* __restrict__ sn,
                                                           int cs, float* __restrict__ noise, int Lf, long long total) {
    TVC_PDL_PROLOGUE();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int L = Lf * kFrame;
    const long long b = i / L;
    const int n = (int)(i - b * L);
    const int j = n / kFrame, r = n - j * kFrame;
    const int tlo = j - 1 < 0 ? 0 : j - 1;
    const int thi = j + 2 > Lf ? Lf : j + 2;
    float acc = 0.f;
    for (int t = tlo; t <= thi; ++t) {
        if (t == 0) continue;                      // the prepended zero frame only counts in the envelope
        const int q = kFrame * (j + 2 - t) + r;    // position inside frame t, 0..1919
        const long long row = b * Lf + (t - 1);
        float v;
        if (q <= kNfft / 2) v = __fsub_rn(__ldg(c + row * cs + q), __ldg(sn + row * cs + q));
        else v = __fadd_rn(__ldg(c + row * cs + (kNfft - q)), __ldg(sn + row * cs + (kNfft - q)));
        acc = __fadd_rn(acc, v);
    }
    noise[i] = __fdiv_rn(acc, (float)(thi - tlo + 1));
}

// ---------------------------------------------------------------------------------------------
// Harmonic oscillator (decoder.py:24-54).  For utterance b, oscillator k = 1..15,
// sample n:  inc = fp32(fp32(interp(f0)[n] * k) / 24000);  I = fp32( sum_{m<=n} fp64(inc[m]) );
// theta = fp32(2 pi) * fmodf(I, 1);  h = sinf(theta) * interp(f0 > 20)[n];  src = h * interp(amps)[n].
// torch.cumsum (CPU) accumulates the fp32 increments in fp64 and rounds every prefix to fp32, so the
// scan must be carried in fp64.  The fp64 sums are exact (or off in bit 53, far below the fp32
// rounding) in any association, so the scan is done in three levels: per-frame totals (kernel 1),
// an exclusive scan of the frame totals per (b,k) (kernel 2), and sequential runs + one warp scan inside
// each frame (kernel 3), which also evaluates the oscillators, multiplies the interpolated amplitudes and
// writes the FilterNet's 17-channel input (15 harmonics, noise, energy; decoder.py:224,265) as
// channels-last split planes.
// ---------------------------------------------------------------------------------------------
// Thread layout shared by the two per-frame kernels: block = one frame (480 samples) of one utterance,
// warp k = oscillator k+1, lane l = the 15 consecutive samples [15 l, 15 l + 15) of the frame, summed
// sequentially in fp64; one warp scan over the 32 lane totals then gives every sample's prefix.
constexpr int kRun = kFrame / 32;   // 15 samples per lane
static_assert(kRun * 32 == kFrame, "frame must split into 32 equal runs");

// The interpolated f0 of a sample is shared by the 15 oscillators (and by the two passes over the increments), so
// each block first evaluates it once per sample into shared memory; an increment is then fp32(fp32(fsi * k) / 24000).
__device__ __forceinline__ float osc_f0_at(const float* __restrict__ f0b, int n, float scale_size, int Lf) {
    const LinCoord c = lin_coord(n, scale_size, Lf);
    return lin_blend(__ldg(f0b + c.i0), __ldg(f0b + c.i1), c);
}
__device__ __forceinline__ float osc_increment(float fsi, float kf) { return __fdiv_rn(__fmul_rn(fsi, kf), kSampleRate); }

__global__ void __launch_bounds__(kOsc * 32) osc_frame_sums_kernel(const float* __restrict__ f0, double* __restrict__ totals,
                                                                   int Lf, float scale_size) {
    TVC_PDL_PROLOGUE();
    __shared__ float s_fsi[kFrame];
    const int fr = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const float* f0b = f0 + (long long)b * Lf;
    s_fsi[threadIdx.x] = osc_f0_at(f0b, fr * kFrame + threadIdx.x, scale_size, Lf);
    __syncthreads();
    const float* run = s_fsi + lane * kRun;       // stride 15 words between lanes: conflict-free
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < kRun; ++i) v = __dadd_rn(v, (double)osc_increment(run[i], (float)(k + 1)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) totals[((long long)b * Lf + fr) * kOsc + k] = v;
}

// in place: totals[b][fr][k] -> sum of totals[b][0..fr-1][k]
__global__ void __launch_bounds__(kOsc * 32) osc_scan_frames_kernel(double* __restrict__ totals, int Lf) {
    TVC_PDL_PROLOGUE();
    const int b = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* tb = totals + (long long)b * Lf * kOsc + k;
    double carry = 0.0;
    for (int base = 0; base < Lf; base += 32) {
        const int fr = base + lane;
        const double own = fr < Lf ? tb[(long long)fr * kOsc] : 0.0;
        double v = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = __dadd_rn(v, u);
        }
        if (fr < Lf) tb[(long long)fr * kOsc] = __dadd_rn(carry, __dadd_rn(v, -own));
        carry = __dadd_rn(carry, __shfl_sync(0xffffffffu, v, 31));
    }
}

__global__ void __launch_bounds__(kOsc * 32, 2) osc_source_kernel(const float* __restrict__ f0, const double* __restrict__ carry,
                                                               const float* __restrict__ amps, int amps_cs,
                                                               const float* __restrict__ noise,
                                                               const float* __restrict__ energy, bf16* __restrict__ src_hi,
                                                               bf16* __restrict__ src_lo, int src_cs, int Lf,
                                                               float scale_size, float scale_factor) {
    TVC_PDL_PROLOGUE();
    __shared__ float tile[kFrame][kOsc + 2];      // [sample][oscillator], 17-float rows: conflict-free both ways
    __shared__ float s_fsi[kFrame], s_uv[kFrame], s_al0[kFrame], s_al1[kFrame];
    __shared__ int s_ai0[kFrame];
    const int fr = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const float* f0b = f0 + (long long)b * Lf;
    {
        // per-sample quantities shared by all oscillators: interp(f0), interp(f0 > 20), amplitude interpolation coords
        const int n = fr * kFrame + threadIdx.x;
        const LinCoord c = lin_coord(n, scale_size, Lf);
        const float fa = __ldg(f0b + c.i0), fb = __ldg(f0b + c.i1);
        s_fsi[threadIdx.x] = lin_blend(fa, fb, c);
        s_uv[threadIdx.x] = lin_blend(fa > 20.0f ? 1.f : 0.f, fb > 20.0f ? 1.f : 0.f, c);
        const LinCoord ca = lin_coord(n, scale_factor, Lf);
        s_ai0[threadIdx.x] = ca.i0 | (ca.i1 != ca.i0 ? 0x40000000 : 0);
        s_al0[threadIdx.x] = ca.l0;
        s_al1[threadIdx.x] = ca.l1;
    }
    __syncthreads();
    const int r0 = lane * kRun;
    const float kf = (float)(k + 1);
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < kRun; ++i) v = __dadd_rn(v, (double)osc_increment(s_fsi[r0 + i], kf));
    double ex = v;                                  // exclusive scan of the lane totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, ex, o);
        if (lane >= o) ex = __dadd_rn(ex, u);
    }
    // second pass: the same increments again (cheaper than keeping 15 fp64 prefixes live), now accumulated
    // on top of everything that precedes this lane's run.  (base + run prefix) in one chain: the sums are
    // exact in fp64, so the association does not change the fp32 rounding of I.
    double acc = __dadd_rn(carry[((long long)b * Lf + fr) * kOsc + k], __dadd_rn(ex, -v));
    const float* ab = amps + (long long)b * Lf * amps_cs + k;
#pragma unroll 5
    for (int i = 0; i < kRun; ++i) {
        acc = __dadd_rn(acc, (double)osc_increment(s_fsi[r0 + i], kf));
        const float I = __double2float_rn(acc);
        // I % 1 (decoder.py:50): for I >= 0 fmodf(I, 1) == I - floorf(I), and that difference is exact in fp32
        const float theta = __fmul_rn(6.28318530717958647692f, I >= 0.f ? __fsub_rn(I, floorf(I)) : fmodf(I, 1.0f));
        const float h = __fmul_rn(sinf(theta), s_uv[r0 + i]);
        const int ai = s_ai0[r0 + i];
        const int i0 = ai & 0x3fffffff, i1 = i0 + (ai >> 30);
        const float a = __fmaf_rn(__ldg(ab + (long long)i0 * amps_cs), s_al0[r0 + i], __fmul_rn(__ldg(ab + (long long)i1 * amps_cs), s_al1[r0 + i]));
        tile[r0 + i][k] = __fmul_rn(h, a);
    }
    __syncthreads();
    // one thread per sample assembles the 24-channel row: 15 harmonics, noise, energy, 7 zeros (decoder.py:224,265)
    const int n = threadIdx.x;
    const long long row = ((long long)b * Lf + fr) * kFrame + n;
    float out[24];
#pragma unroll
    for (int q = 0; q < kOsc; ++q) out[q] = tile[n][q];
    out[15] = __ldg(noise + row);
    out[16] = __ldg(energy + row);
#pragma unroll
    for (int q = 17; q < 24; ++q) out[q] = 0.f;
    uint32_t hh[12], ll[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        bf16 h0, l0, h1, l1;
        split_bf16(out[2 * q], h0, l0);
        split_bf16(out[2 * q + 1], h1, l1);
        hh[q] = pack2(h0, h1);
        ll[q] = pack2(l0, l1);
    }
    uint4* ph = reinterpret_cast<uint4*>(src_hi + row * src_cs);
    uint4* pl = reinterpret_cast<uint4*>(src_lo + row * src_cs);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        ph[q] = make_uint4(hh[4 * q], hh[4 * q + 1], hh[4 * q + 2], hh[4 * q + 3]);
        pl[q] = make_uint4(ll[4 * q], ll[4 * q + 1], ll[4 * q + 2], ll[4 * q + 3]);
    }
}

}  // namespace

int noise_spectrum_cl(const float* kern, int k_cs, const float* rand01, bf16* yr_hi, bf16* yr_lo, bf16* yi_hi,
                      bf16* yi_lo, int y_cs, int B, int Lf, cudaStream_t s) {
    dim3 grid(cdiv(Lf, 32), cdiv(y_cs, 32), B);
    TVC_LAUNCH_PDL(noise_spectrum_cl_kernel, grid, 256, 0, s, kern, k_cs, rand01, yr_hi, yr_lo, yi_hi, yi_lo, y_cs, Lf);
    TVC_LAUNCH_CHECK();
    return 0;
}

int noise_ola_cl(const float* c, const float* sn, int cs, float* noise, int B, int Lf, cudaStream_t s) {
    const long long total = (long long)B * Lf * kFrame;
    TVC_LAUNCH_PDL(noise_ola_cl_kernel, cdiv(total, 256), 256, 0, s, c, sn, cs, noise, Lf, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

size_t osc_scratch_bytes(int B, int Lf) { return sizeof(double) * (size_t)B * Lf * kOsc; }

int harmonic_source_cl(const float* f0, const float* amps, int amps_cs, const float* noise, const float* energy,
                       bf16* src_hi, bf16* src_lo, int src_cs, void* scratch, int B, int Lf, cudaStream_t s) {
    TVC_REQUIRE(src_cs == 24, "harmonic_source_cl: source planes must have 24 channels");
    const int L = Lf * kFrame;
    const float scale_size = (float)Lf / (float)L;              // F.interpolate(size=L)
    const float scale_factor = (float)(1.0 / (double)kFrame);   // F.interpolate(scale_factor=480)
    double* totals = (double*)scratch;
    dim3 grid(Lf, B);
    TVC_LAUNCH_PDL(osc_frame_sums_kernel, grid, kOsc * 32, 0, s, f0, totals, Lf, scale_size);
    TVC_LAUNCH_CHECK();
    TVC_LAUNCH_PDL(osc_scan_frames_kernel, B, kOsc * 32, 0, s, totals, Lf);
    TVC_LAUNCH_CHECK();
    TVC_LAUNCH_PDL(osc_source_kernel, grid, kOsc * 32, 0, s, f0, totals, amps, amps_cs, noise, energy, src_hi, src_lo, src_cs, Lf,
                                               scale_size, scale_factor);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
