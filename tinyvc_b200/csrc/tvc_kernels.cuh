// Internal kernel-launcher declarations (one per hot-path op).  All pointers are device pointers.
#pragma once
#include "tvc_common.cuh"

namespace tvc {

// frame_ops.cu
int frame_prep(const float* energy, const float* f0, float* e_fr, float* lf0, int B, int Lf, cudaStream_t s);
int rank1_add(float* x, const float* a1, const float* w1, const float* b1, const float* a2, const float* w2,
              const float* b2, int B, int C, int T, cudaStream_t s);
int dwconv_ln(const float* x, float* y, const float* w, const float* wb, const float* gamma, const float* beta,
              int B, int C, int T, int dil, cudaStream_t s);
int grn_scale(const float* y, const float* gamma, float* scale, int B, int C, int T, cudaStream_t s);
int pitch_decode(const float* logits, float* f0, int B, int ncls, int T, cudaStream_t s);
int interp_linear(const float* x, float* y, long long rows, int tin, int tout, float scale, cudaStream_t s);
int out_conv_k7(const float* x, const float* w, const float* bias, float* y, int B, int C, int T, cudaStream_t s);
int repack_conv_weight(const float* src, float* dst, int Cout, int Cin, int K, int CoutP, int co_off, cudaStream_t s);

// dsp.cu
int harmonic_osc(const float* f0, const float* amps, float* src, long long src_bs, int B, int Lf, cudaStream_t s);
int harmonic_theta(const float* f0, float* theta, int B, int Lf, cudaStream_t s);
int noise_spectrum(const float* kern, const float* rand01, float* yri, int B, int Lf, cudaStream_t s);
int noise_ola(const float* cs, float* src, long long src_bs, int ch, int B, int Lf, cudaStream_t s);

// frontend.cu
int energy_pooled_len(int L);
int energy_estimate(const float* wf, float* energy, float* pooled, int B, int L, cudaStream_t s);
int shift_frequency(const float* f0, float* out, long long n, float shift, cudaStream_t s);

// resample.cu
long long resample_length(long long L, int orig, int neu);
int resample_run(const float* x, float* y, int B, long long L, int orig, int neu, cudaStream_t s);

// knn.cu
int knn_prepare(const float* index_cn, float* index_w, float* index_nc, float* bias, int C, int N, int NP, int metric,
                cudaStream_t s);
int knn_normalize_queries(const float* src, float* qn, int B, int C, int T, int metric, cudaStream_t s);
int knn_topk(const float* sims, float* pv, int* pi, int* idx_out, int B, int T, int N, int k, cudaStream_t s);
int knn_gather_mean(const float* src, const float* index_nc, const int* idx, float* out, int B, int C, int T, int k,
                    float alpha, cudaStream_t s);
int knn_rescore_candidates(const float* qn, const float* index_wn, const int* cand, const float* score, int splits, int* flag,
                           int* idx_out, int B, int T, int N, int k, cudaStream_t s);
constexpr float kKnnScreenEps = 1e-4f;   // screening margin: >= 2 x the worst-case error of the split-bf16 similarity product
int knn_transpose_index(const float* index_w, float* index_wn, int C, int N, int NP, cudaStream_t s);
constexpr int kKnnMaxKDecl = 8;
constexpr int kKnnSegDecl = 16;

// stream.cu
int sola_run(const float* y, int y_len, float* sola_buf, const float* fade_in, float* out_block, int* shift_out, int S,
             int block, int cross, int search, int delay, cudaStream_t s, float* pv_scratch = nullptr);
size_t phase_vocoder_scratch_floats(int S, int n);
int phase_vocoder_run(const float* ab, long long a_stride, long long b_off, const float* fade_in, float* bins, float* out,
                      long long out_stride, int S, int n, cudaStream_t s);

}  // namespace tvc
