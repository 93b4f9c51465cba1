// Tensor-core (tcgen05 / TMEM) dense Conv1d for the TinyVC decoder -- declarations.
//
// Activations on this path are channels-last "split planes": a tensor [rows = B*T][C] is stored
// as two bf16 matrices hi = bf16(v), lo = bf16(v - hi) with `cs` channels of capacity (a multiple of
// 8 elements = one 16-byte UMMA core-matrix row; channels [C, cs) hold finite padding, their
// weights are zero).
//
// Memory order is CHUNK-MAJOR: element (row, ch) of a tensor with R rows lives at
//     cm(row, ch, R) = ((ch / 8) * R + row) * 8 + ch % 8
// i.e. every 8-channel chunk is its own dense [R][8] array (16 bytes per row for bf16, 32 for fp32).  That
// is the order the tensor cores consume (the smem operand of one K-chunk is a column of rows at a 16-byte
// pitch): a warp moving 32 consecutive rows of one chunk touches 512 contiguous bytes on both the global
// and the shared side, for the conv kernel's operand gathers and for its epilogue stores alike.  fp32
// companions (residuals, resampler inputs) use the same order.  A view on channels [c0, c0 + n) with
// c0 % 8 == 0 of a tensor with the same R is just a pointer offset of cm(0, c0, R).  A conv is evaluated as three bf16 tensor-core products accumulated in
// fp32 in TMEM:  x*w ~= x_hi*w_hi + x_hi*w_lo + x_lo*w_hi  (relative error ~2^-16, against
// 2^-11 for one TF32 product, which SURVEY.md section 0 shows is not enough for RMSE < 1e-4).
#pragma once
#include <cuda_bf16.h>

#include "tvc_common.cuh"

namespace tvc {

typedef __nv_bfloat16 bf16;

__host__ __device__ __forceinline__ long long cm(long long row, int ch, long long R) {
    return ((long long)(ch >> 3) * R + row) * 8 + (ch & 7);
}

enum TcAux { TC_AUX_NONE = 0, TC_AUX_ACC = 1, TC_AUX_FILM = 2 };
enum TcAct { TC_ACT_NONE = 0, TC_ACT_LRELU = 1, TC_ACT_GELU = 2, TC_ACT_ELU1 = 3 };

// Packed weights of one conv (device memory, built once at load by tc_pack_conv).
struct TcConvW {
    bf16* w = nullptr;           // per n-tile, per K-stage: [hi image][lo image] in UMMA K-major core-matrix order
    float* bias = nullptr;       // [n_tiles * NTp]   (main bias, + aux bias for TC_AUX_ACC)
    float* film_bias = nullptr;  // [n_tiles * 2*NTp] (scale bias | shift bias), TC_AUX_FILM only
    int Cin = 0, Cout = 0, taps = 1;
    int aux_cin = 0, aux_mode = TC_AUX_NONE;
    int KB = 0, nkb = 0, aux_nkb = 0;   // K-stage = KB channels of one tap
    int NT = 0, NTp = 0, n_tiles = 0;   // output channels per CTA, padded to 16
    // "cat" images (the fused block kernel's, tc_block.cu): in the main K-stages the hi and lo rows of a chunk are adjacent
    // ([chunk][hi NTp | lo NTp][8]), so x_hi * [w_hi | w_lo] is ONE MMA of N = 2 NTp and a K-step costs two instructions
    // instead of three (tc_ptx.cuh issue_stage_cat).  tc_conv_kernel itself takes plain images only.
    bool cat = false;
    size_t tile_elems = 0;              // bf16 elements of packed weights per n-tile
    void free_all();
};

// `w` host [Cout][Cin][taps] (torch Conv1d layout), `b` host [Cout] (nullable).
// aux (1x1 on a second input, host pointers):
//   TC_AUX_ACC : aux_w [Cout][aux_cin], aux_b [Cout]  -> accumulated into the same output
//   TC_AUX_FILM: aux_w [2][Cout][aux_cin] (scale rows then shift rows), aux_b [2][Cout]
//                -> out = conv * (scale) + shift (+ res)            (decoder.py:88-97)
int tc_pack_conv(const float* w, const float* b, int Cout, int Cin, int taps, const float* aux_w, const float* aux_b,
                 int aux_cin, int aux_mode, int NT, TcConvW& out, bool cat = false);

struct TcConvArgs {
    const bf16 *a_hi = nullptr, *a_lo = nullptr;   // main input planes, chunk-major, a_cs channels of capacity
    int a_cs = 0;
    const bf16 *x_hi = nullptr, *x_lo = nullptr;   // aux input planes (same rows)
    int x_cs = 0;
    int dil = 1;
    int B = 0, T = 0;                  // taps are clamped inside each utterance of T rows (replicate padding)
    // Stored replicate padding ("padded mode", k = 3 convs): the main input planes hold a_pad >= dil extra rows on either
    // side of every utterance (rows b * (T + 2 * a_pad) + [0, a_pad) repeat t = 0, the last a_pad repeat t = T - 1), so a
    // tap is a shifted view and every operand window is one contiguous tensor copy, whatever T is (short utterances
    // otherwise gather their clamped windows row by row).  y_pad: the plane output is written in the same form for
    // the next k = 3 conv (its replicate rows included).  The aux input, the residual and y32 are never padded.
    int a_pad = 0, y_pad = 0;
    const float* res = nullptr;        // fp32 residual (chunk-major, same rows), added after bias / FiLM
    int res_cs = 0;
    float* y32 = nullptr;              // fp32 output (chunk-major, nullable)
    int y32_cs = 0;
    bf16 *y_hi = nullptr, *y_lo = nullptr;   // split-plane output (chunk-major, nullable)
    int y_cs = 0;
    // Fused top-k (kNN screening): with topk_cand set the conv writes no tensor; per row it keeps the 8 largest of columns
    // [0, topk_n) (ties: lower index) and stores their indices [rows][8] plus topk_flag[row] = 1 when the topk_k-th and the
    // 8-th scores are closer than topk_eps.  1 x 1 convs without aux / residual only.
    int* topk_cand = nullptr;
    int* topk_flag = nullptr;
    int topk_k = 0, topk_n = 0;
    float topk_eps = 0.f;
    // Split sweep (few query tiles, e.g. a streaming tick: 28 row tiles on 148 SMs): the channel tiles are cut into
    // topk_splits equal ranges and (row tile, range) pairs are spread over the CTAs; every pair stores its own 8 candidates
    // AND their scores -- topk_cand / topk_score are then [rows][topk_splits][8], topk_flag is not written, and the caller
    // merges the lists (knn.cu: the merged list is exactly the single sweep's).  tc_conv_topk_splits() picks the count.
    float* topk_score = nullptr;
    int topk_splits = 1;
    // Fused down-resampler: with dec_f in {3, 4, 5} the epilogue also writes F.interpolate(y, scale_factor = 1 / dec_f,
    // mode='linear') of the conv's result (interp_cl's arithmetic) as the next Downsample block's two operands: raw planes
    // dec_r (B * T / dec_f rows) and leaky-ReLU'd planes dec_a (with dec_pad stored replicate rows per utterance side).
    // Needs a plane output and no fp32 output / residual / FiLM / activation.
    int dec_f = 0, dec_pad = 0;
    float dec_scale = 0.f;             // interp_cl's scale argument for this factor
    bf16 *dec_r_hi = nullptr, *dec_r_lo = nullptr, *dec_a_hi = nullptr, *dec_a_lo = nullptr;
    int epi_act = TC_ACT_NONE;         // applied to the value (both outputs)
    int out_act = TC_ACT_NONE;         // applied additionally to the split-plane copy only
};
int tc_conv_launch(const TcConvW& W, const TcConvArgs& a, cudaStream_t s);
int tc_conv_topk_splits(const TcConvW& W, long long rows);   // TcConvArgs::topk_splits for this many query rows
int tc_conv_init();

// developer timeline of selected tc_conv launches (ordinals counted from arming); see tc_conv.cu Tracer
int tc_trace_arm(const char* ordinals);
int tc_trace_dump(const char* path);

int measure_fp32_peak(double* tflops, cudaStream_t s);   // tc_ops.cu: FMA micro-benchmark

// layout helpers (tc_ops.cu)
// extra0/extra1 (nullable, [B*T]) are appended as channels C and C+1 (frame-rate scalars such as log-f0).
int cf_to_planes(const float* x, bf16* hi, bf16* lo, int B, int C, int T, int cs, int act, cudaStream_t s,
                 const float* extra0 = nullptr, const float* extra1 = nullptr);
int cf_to_cl(const float* x, float* y, int B, int C, int T, int cs, cudaStream_t s);
int cl_to_cf(const float* x, float* y, int B, int C, int T, int cs, cudaStream_t s);
int planes_to_cf(const bf16* hi, const bf16* lo, float* y, int B, int C, int T, int cs, cudaStream_t s);
// re-pads one plane: out row (b, tp) of a tensor with pad_out replicate rows per side <- in row (b, clamp(tp - pad_out, 0, T - 1))
// of a tensor with pad_in (parity probes: builds padded inputs, strips padded outputs)
int plane_repad(const bf16* in, bf16* out, int B, int T, int cs, int pad_in, int pad_out, cudaStream_t s);

}  // namespace tvc
