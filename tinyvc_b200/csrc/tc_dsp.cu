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
// Counter-based uniform generator for the noise phases when the caller does not inject the draw:
// Philox-4x32-10 keyed by the decoder's seed; counter = (element group, launch step).  One call yields
// four 32-bit words = four uniforms u = (w >> 8) * 2^-24 in [0, 1), the grid torch.rand uses on the CPU.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }

__global__ void rng_seed_kernel(unsigned long long* state, unsigned long long seed) {
    state[0] = seed;
    state[1] = 0ull;
}
__global__ void rng_advance_kernel(unsigned long long* state) {
    TVC_PDL_PROLOGUE();
    state[1] += 1ull;
}

// ---------------------------------------------------------------------------------------------
// noise_spectrum_cl (decoder.py:78-80): angle = (rand01*2)*pi - pi, Y = kernel * exp(j*angle).
// rand01 arrives channels-first [B][961][Lf] (the layout torch.rand draws it in); a 32x32 tile
// transpose through shared memory turns it into channels-last rows.  Outputs: split planes of
// Re(Y) and Im(Y) (chunk-major, B*Lf rows, y_cs channels of capacity), channels >= 961 zero.
// `kern` is the fp32 head output (chunk-major, same rows; its first 961 channels are the noise filter).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) noise_spectrum_cl_kernel(const float* __restrict__ kern,
                                                                const float* __restrict__ rand01,
                                                                const unsigned long long* __restrict__ rng_state,
                                                                bf16* __restrict__ yr_hi, bf16* __restrict__ yr_lo,
                                                                bf16* __restrict__ yi_hi, bf16* __restrict__ yi_lo,
                                                                int y_cs, int Lf) {
    TVC_PDL_PROLOGUE();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long R = (long long)gridDim.z * Lf;
    if (rand01) {
        for (int j = ty; j < 32; j += 8) {
            const int c = c0 + j, t = t0 + tx;
            tile[j][tx] = (c < kBins && t < Lf) ? __ldg(rand01 + ((long long)b * kBins + c) * Lf + t) : 0.f;
        }
    }
    __syncthreads();
    const float pi_f = 3.14159265358979323846f;
    if (threadIdx.x >= 128) return;
    // one (frame, 8-bin chunk) per thread: 8 bins of Re and Im, stored as one 16-byte row of each chunk array
    const int q = threadIdx.x >> 5, t = t0 + tx, cq = c0 + q * 8;
    if (t >= Lf || cq >= y_cs) return;
    const long long row = (long long)b * Lf + t;
    const long long o = cm(row, cq, R);
    const float4 k0 = __ldg(reinterpret_cast<const float4*>(kern + o)), k1 = __ldg(reinterpret_cast<const float4*>(kern + o) + 1);
    const float kv[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    float u[8];
    if (rand01) {
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] = tile[q * 8 + e][tx];
    } else {
        // torch.rand(B, 961, Lf) of decoder.py:78, drawn here: element (row, bin) <- word (bin % 4) of block (row * 242 + bin / 4)
        const unsigned long long seed = rng_state[0], step = rng_state[1];
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        const unsigned long long blk = (unsigned long long)row * 242ull + (unsigned long long)(cq >> 2);
        const uint4 r0 = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)step, (uint32_t)(step >> 32)), key);
        const uint4 r1 = philox4x32_10(make_uint4((uint32_t)(blk + 1), (uint32_t)((blk + 1) >> 32), (uint32_t)step, (uint32_t)(step >> 32)), key);
        u[0] = u01(r0.x); u[1] = u01(r0.y); u[2] = u01(r0.z); u[3] = u01(r0.w);
        u[4] = u01(r1.x); u[5] = u01(r1.y); u[6] = u01(r1.z); u[7] = u01(r1.w);
    }
    float re[8], im[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        re[e] = 0.f; im[e] = 0.f;
        if (cq + e < kBins) {
            const float a = __fsub_rn(__fmul_rn(__fmul_rn(u[e], 2.0f), pi_f), pi_f);
            float sn, cs;
            sincosf(a, &sn, &cs);
            re[e] = __fmul_rn(cs, kv[e]);
            im[e] = __fmul_rn(sn, kv[e]);
        }
    }
    uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        bf16 h0, l0, h1, l1;
        split_bf16(re[2 * e], h0, l0); split_bf16(re[2 * e + 1], h1, l1);
        rh[e] = pack2(h0, h1); rl[e] = pack2(l0, l1);
        split_bf16(im[2 * e], h0, l0); split_bf16(im[2 * e + 1], h1, l1);
        ih[e] = pack2(h0, h1); il[e] = pack2(l0, l1);
    }
    *reinterpret_cast<uint4*>(yr_hi + o) = make_uint4(rh[0], rh[1], rh[2], rh[3]);
    *reinterpret_cast<uint4*>(yr_lo + o) = make_uint4(rl[0], rl[1], rl[2], rl[3]);
    *reinterpret_cast<uint4*>(yi_hi + o) = make_uint4(ih[0], ih[1], ih[2], ih[3]);
    *reinterpret_cast<uint4*>(yi_lo + o) = make_uint4(il[0], il[1], il[2], il[3]);
}

// ---------------------------------------------------------------------------------------------
// noise_ola_cl (decoder.py:81-82 = torch.istft, n_fft 1920, hop 480, rectangular window, centred,
// one all-zero frame prepended).  The inverse real DFT of frame t is x[p] = C[p] - S[p],
// x[N-p] = C[p] + S[p] (p <= N/2) with C = cos-basis * Re(Y), S = sin-basis * Im(Y) computed by
// two tensor-core products; this kernel overlap-adds and divides by the window coverage.
// c, sn: fp32 chunk-major with B*Lf rows;  noise: [B][L].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) noise_ola_cl_kernel(const float* __restrict__ c, const float* __restrict__ sn,
                                                           long long R, float* __restrict__ noise, int Lf, long long total) {
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
        if (q <= kNfft / 2) v = __fsub_rn(__ldg(c + cm(row, q, R)), __ldg(sn + cm(row, q, R)));
        else v = __fadd_rn(__ldg(c + cm(row, kNfft - q, R)), __ldg(sn + cm(row, kNfft - q, R)));
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
                                                               const float* __restrict__ amps,
                                                               const float* __restrict__ noise,
                                                               const float* __restrict__ energy, bf16* __restrict__ src_hi,
                                                               bf16* __restrict__ src_lo, int Lf,
                                                               float scale_size, float scale_factor) {
    TVC_PDL_PROLOGUE();
    __shared__ float tile[kFrame][kOsc + 2];      // [sample][oscillator], 17-float rows: conflict-free both ways
    __shared__ float s_fsi[kFrame], s_uv[kFrame], s_al0[kFrame], s_al1[kFrame];
    __shared__ int s_ai0[kFrame];
    __shared__ float s_amp[3][kOsc + 1];          // amplitudes of frames fr - 1, fr, fr + 1 (all a frame's samples interpolate between)
    const int fr = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    const float* f0b = f0 + (long long)b * Lf;
    if (threadIdx.x < 3 * kOsc) {
        const int w = threadIdx.x / kOsc, q = threadIdx.x - w * kOsc;
        int f = fr - 1 + w;
        f = f < 0 ? 0 : (f > Lf - 1 ? Lf - 1 : f);
        const long long RF0 = (long long)gridDim.y * Lf;
        s_amp[w][q] = __ldg(amps + cm((long long)b * Lf + f, q, RF0));
    }
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
    const long long RF = (long long)gridDim.y * Lf;           // rows of the frame-rate tensors
#pragma unroll 5
    for (int i = 0; i < kRun; ++i) {
        acc = __dadd_rn(acc, (double)osc_increment(s_fsi[r0 + i], kf));
        const float I = __double2float_rn(acc);
        // I % 1 (decoder.py:50): for I >= 0 fmodf(I, 1) == I - floorf(I), and that difference is exact in fp32
        const float theta = __fmul_rn(6.28318530717958647692f, I >= 0.f ? __fsub_rn(I, floorf(I)) : fmodf(I, 1.0f));
        const float h = __fmul_rn(sinf(theta), s_uv[r0 + i]);
        const int ai = s_ai0[r0 + i];
        const int i0 = ai & 0x3fffffff, i1 = i0 + (ai >> 30);
        // i0, i1 lie in [fr - 1, fr + 1] (clamped to the utterance exactly like lin_coord clamps them)
        const float a = __fmaf_rn(s_amp[i0 - fr + 1][k], s_al0[r0 + i], __fmul_rn(s_amp[i1 - fr + 1][k], s_al1[r0 + i]));
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
    const long long RL = RF * kFrame;                           // rows of the sample-rate tensors
#pragma unroll
    for (int q = 0; q < 3; ++q) {                                // chunk-major: 32 consecutive samples of a chunk = 512 B
        *reinterpret_cast<uint4*>(src_hi + (q * RL + row) * 8) = make_uint4(hh[4 * q], hh[4 * q + 1], hh[4 * q + 2], hh[4 * q + 3]);
        *reinterpret_cast<uint4*>(src_lo + (q * RL + row) * 8) = make_uint4(ll[4 * q], ll[4 * q + 1], ll[4 * q + 2], ll[4 * q + 3]);
    }
}

}  // namespace

int noise_spectrum_cl(const float* kern, const float* rand01, unsigned long long* rng_state, bf16* yr_hi, bf16* yr_lo,
                      bf16* yi_hi, bf16* yi_lo, int y_cs, int B, int Lf, cudaStream_t s) {
    TVC_REQUIRE(rand01 || rng_state, "noise_spectrum: neither an injected draw nor a generator state");
    dim3 grid(cdiv(Lf, 32), cdiv(y_cs, 32), B);
    TVC_LAUNCH_PDL(noise_spectrum_cl_kernel, grid, 256, 0, s, kern, rand01, (const unsigned long long*)rng_state, yr_hi, yr_lo,
                   yi_hi, yi_lo, y_cs, Lf);
    TVC_LAUNCH_CHECK();
    if (!rand01) {                       // the next call draws a fresh tensor
        TVC_LAUNCH_PDL(rng_advance_kernel, 1, 1, 0, s, rng_state);
        TVC_LAUNCH_CHECK();
    }
    return 0;
}

int rng_seed(unsigned long long* state, unsigned long long seed, cudaStream_t s) {
    rng_seed_kernel<<<1, 1, 0, s>>>(state, seed);
    TVC_LAUNCH_CHECK();
    return 0;
}

int noise_ola_cl(const float* c, const float* sn, float* noise, int B, int Lf, cudaStream_t s) {
    const long long total = (long long)B * Lf * kFrame;
    TVC_LAUNCH_PDL(noise_ola_cl_kernel, cdiv(total, 256), 256, 0, s, c, sn, (long long)B * Lf, noise, Lf, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

size_t osc_scratch_bytes(int B, int Lf) { return sizeof(double) * (size_t)B * Lf * kOsc; }

// the first two levels of the oscillator's scan (per-frame totals, per-utterance frame scan): they depend on f0 only, so the
// plan may run them early on a side branch (harmonic_source_cl(..., scan_done = true) then starts at the third level)
int osc_phase_scan(const float* f0, void* scratch, int B, int Lf, cudaStream_t s) {
    const int L = Lf * kFrame;
    const float scale_size = (float)Lf / (float)L;              // F.interpolate(size=L)
    double* totals = (double*)scratch;
    dim3 grid(Lf, B);
    TVC_LAUNCH_PDL(osc_frame_sums_kernel, grid, kOsc * 32, 0, s, f0, totals, Lf, scale_size);
    TVC_LAUNCH_CHECK();
    TVC_LAUNCH_PDL(osc_scan_frames_kernel, B, kOsc * 32, 0, s, totals, Lf);
    TVC_LAUNCH_CHECK();
    return 0;
}

int harmonic_source_cl(const float* f0, const float* amps, const float* noise, const float* energy,
                       bf16* src_hi, bf16* src_lo, void* scratch, int B, int Lf, cudaStream_t s, bool scan_done) {
    const int L = Lf * kFrame;
    const float scale_size = (float)Lf / (float)L;              // F.interpolate(size=L)
    const float scale_factor = (float)(1.0 / (double)kFrame);   // F.interpolate(scale_factor=480)
    double* totals = (double*)scratch;
    dim3 grid(Lf, B);
    if (!scan_done) TVC_TRY(osc_phase_scan(f0, scratch, B, Lf, s));
    TVC_LAUNCH_PDL(osc_source_kernel, grid, kOsc * 32, 0, s, f0, totals, amps, noise, energy, src_hi, src_lo, Lf,
                                               scale_size, scale_factor);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc
