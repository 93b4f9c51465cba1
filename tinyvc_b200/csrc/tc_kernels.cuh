// Launcher declarations of the channels-last (tensor-core path) helper kernels.
#pragma once
#include "tc_conv.cuh"

namespace tvc {

// tc_frame.cu
// all channels-last tensors are chunk-major (tc_conv.cuh); row counts follow from B and T
// a_pad > 0: the activation planes (a_hi / a_lo) are written with stored replicate padding (tc_conv.cuh, padded mode)
// windowed form (output pruning): the output holds rows [out_off, out_off + Tout) of the full-length result, the input rows
// [in_off, in_off + Tin_c) of the full-length (Tin rows) input; Tin_c <= 0: the plain full-length call
int interp_cl(const float* x, int B, int Tin, int Tout, float scale, int C, float* y32, bf16* r_hi, bf16* r_lo, bf16* a_hi,
              bf16* a_lo, cudaStream_t s, int a_pad = 0, int Tin_c = 0, int in_off = 0, int out_off = 0);
// rows [t0, t0 + Tc) of every utterance of a plane pair (B * T rows, C channels of capacity) -> compact planes (B * Tc rows)
int slice_planes_cl(const bf16* s_hi, const bf16* s_lo, bf16* d_hi, bf16* d_lo, int B, int T, int C, int t0, int Tc, cudaStream_t s);
int dwconv_ln_cl(const float* x, const float* w7, const float* wb, const float* gamma, const float* beta,
                 bf16* hi, bf16* lo, int B, int T, cudaStream_t s);
// general ConvNeXt front half: depth-wise k = 7 (dilation `dil`; w7 [7][C] repacked, nullptr = none) + LayerNorm over C
// (128 or 384) channels -> split planes and / or fp32, chunk-major
int cnxt_ln_cl(const float* x, const float* w7, const float* wb, const float* gamma, const float* beta, bf16* hi, bf16* lo,
               float* y32, int B, int C, int T, int dil, cudaStream_t s);
int grn_apply_cl(const float* y, const float* gamma, const float* beta, bf16* hi, bf16* lo, int B, int C, int T,
                 cudaStream_t s);
int out_conv_k7_cl(const float* x, const float* w, const float* bias, float* y, int B, int T, cudaStream_t s);

// tc_dsp.cu
// rand01 == nullptr: the uniform draw of decoder.py:78 comes from the Philox state {seed, step} at rng_state
int noise_spectrum_cl(const float* kern, const float* rand01, unsigned long long* rng_state, bf16* yr_hi, bf16* yr_lo,
                      bf16* yi_hi, bf16* yi_lo, int y_cs, int B, int Lf, cudaStream_t s);
int rng_seed(unsigned long long* state, unsigned long long seed, cudaStream_t s);
int noise_ola_cl(const float* c, const float* sn, float* noise, int B, int Lf, cudaStream_t s);
size_t osc_scratch_bytes(int B, int Lf);
// amps: view on the 15 amplitude channels (chunk-major, B*Lf rows); src planes: 24 channels of capacity, B*L rows
int harmonic_source_cl(const float* f0, const float* amps, const float* noise, const float* energy,
                       bf16* src_hi, bf16* src_lo, void* scratch, int B, int Lf, cudaStream_t s, bool scan_done = false);
// levels one and two of the oscillator's phase scan alone (f0 only): may run ahead on another stream
int osc_phase_scan(const float* f0, void* scratch, int B, int Lf, cudaStream_t s);

}  // namespace tvc
