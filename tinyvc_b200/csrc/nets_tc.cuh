// Tensor-core (tcgen05) execution plan of the Decoder: packed weights + launch sequence.
#pragma once
#include "nets.cuh"
#include "tc_kernels.cuh"

namespace tvc {

struct DecoderTC {
    TcConvW frame_in;             // [content 768 | e_fr | log f0] -> [SourceNet 128 | FilterNet 384]
    struct Cnxt {
        const float *w7 = nullptr, *wb = nullptr, *ln_g = nullptr, *ln_b = nullptr, *grn_g = nullptr, *grn_b = nullptr;
        TcConvW c2, c3;
    } mid[3];
    TcConvW heads;                // 128 -> [kernel 961 | 7 pad | amps 15]
    TcConvW dft_cos, dft_sin;     // inverse real-DFT bases, 961 -> 961
    TcConvW dft_cos_w, dft_sin_w; // the same with 128-wide channel tiles: on short batches the two products run side by side
    cudaStream_t side = nullptr;  // branch for the sine product (forked from / joined to the caller's stream, also under capture)
    cudaEvent_t fork = nullptr, join = nullptr, scan_ev = nullptr, up0_ev = nullptr;
    TcConvW down0;
    struct Down { TcConvW c1, c2, c3; } down[4];
    struct Up { TcConvW c1, c2, c3, c4, c5; } up[5];
    Up up4_cat;                   // ups.4 again as "cat" images for the fused block kernel
    // Wider channel tiles of the 384- and 192-channel Upsample blocks (96 / 80 wide; FiLM layers need 3 x 2 accumulator columns per
    // output channel in TMEM, so 80 is their limit).  Which image a launch uses is decided per batch: the narrow tiles give a
    // short batch one tile per CTA, the wide ones cut the waves (and the re-streaming of the activation windows) of larger ones
    // -- a streaming tick has 56 row tiles at L/240: 448 narrow tiles = 4 per CTA on 112 CTAs, 224 wide ones = 2 per CTA.
    Up up_w[3];                   // (ups.2: c1 / c3 only -- its FiLM layers cannot be wider than they are)
    const float *out_w = nullptr, *out_b = nullptr;
    float* w7_buf = nullptr;      // depth-wise weights repacked [3][7][128]
    unsigned long long* rng_state = nullptr;   // device {seed, step} of the noise generator (used when no draw is injected)
    bool ready = false;
    ~DecoderTC();
    int init(const WeightStore& store);
    // Decoder.infer (decoder.py:253-257) on the tensor-core path.
    // out_t0 / out_t1: only out[b][out_t0, out_t1) has to be produced (out_t1 < 0: everything); the fused full-rate block
    // then skips the windows outside the range.  Samples outside the range are unspecified.
    int infer(Arena& A, cudaStream_t s, const float* content, const float* f0, const float* energy, const float* rand01,
              float* out, int B, int Lf, int out_t0 = 0, int out_t1 = -1) const;
};

// Tensor-core execution plan of the Encoder (module/tinyvc/encoder.py:11-116): both ConvNeXt stacks with every dense 1x1
// conv on tc_conv_kernel (split bf16 x3, fp32 TMEM accumulation), activations channels-last chunk-major.
struct EncoderTC {
    struct Blk {
        const float *w7 = nullptr, *wb = nullptr, *ln_g = nullptr, *ln_b = nullptr, *grn_g = nullptr, *grn_b = nullptr;
        TcConvW c2, c3;
        TcConvW c2_w;             // c2 again with 192-wide channel tiles (384-channel stack): one wave for a short batch
        int dil = 1;
    };
    struct Stack {
        int C = 0, c0 = 0;            // width, first channel inside the merged input product
        const float *ln_g = nullptr, *ln_b = nullptr;
        std::vector<Blk> mid;
        TcConvW out, out_w;       // output layer (+ its 192-wide variant for the 384-channel stack)
    } ssl, pitch;
    TcConvW in;                   // both input layers as one product over the spectrogram: 961 -> [ssl 384 | pitch 128]
    float* w7_buf = nullptr;      // depth-wise weights repacked [7][C] per block
    bool ready = false;
    ~EncoderTC();
    int init(const WeightStore& store);
    // z [B,768,Lf] and / or logits [B,512,Lf] (channels-first fp32, nullable) from spec [B,961,Lf]
    int forward(Arena& A, cudaStream_t s, const float* spec, float* z, float* logits, int B, int Lf) const;
};

void set_fused_up(bool on);          // tvc_set_option("fused_up", "0"|"1"): fused 24-channel Upsample block (default on)
bool fused_up();
void set_wide_tiles(bool on);        // tvc_set_option("wide_tiles", "0"|"1"): per-batch choice between narrow and wide channel tiles of ups.0 / ups.1 (default on)
void set_side_branch(bool on);       // tvc_set_option("side_branch", "0"|"1"): forked branch inside a decoder step on short batches (default on)
// rows [wa[i], wb[i]) Upsample level i must produce for output samples [out_t0, out_t1) (nets_tc.cu; host arithmetic only)
void decoder_plan_windows(int Lf, int out_t0, int out_t1, bool prune_enabled, int wa[5], int wb[5]);
void set_prune_levels(bool on);      // tvc_set_option("prune_levels", "0"|"1"): output pruning below the fused block (default on)
void set_fuse_down(bool on);         // tvc_set_option("fuse_down", "0"|"1"): Downsample resamplers inside the producing conv's epilogue (default on)
// tvc_set_option("pad_up_max_t" | "pad_down_max_t", "N"): Upsample / Downsample blocks of levels whose utterances have
// <= N rows store the replicate padding of their k = 3 convs (tc_conv.cuh, padded mode); -1 leaves a limit unchanged
void set_pad_max_t(int up, int down);
unsigned plan_options();             // bit set of the options that change the launch plan (part of the graph-cache key)

}  // namespace tvc
