// Tensor-core (tcgen05) execution plan of the Decoder: packed weights + launch sequence.
#pragma once
#include "nets.cuh"
#include "tc_kernels.cuh"

#include <vector>

namespace tvc {

// Where the op tables of chained launches (tc_conv.cuh TcChain) live on the device.  Eager calls carve the region out of the
// workspace and upload with cudaMemcpyAsync; a captured graph owns a device buffer instead (`deferred`): the plan only
// collects the host image and the caller uploads it once, after the capture, so a replay contains no copy at all.
struct ChainSink {
    unsigned char* dev = nullptr;
    size_t cap = 0, used = 0;
    bool deferred = false;
    std::vector<unsigned char> host;
};
constexpr size_t kChainSinkBytes = 96 * 1024;

struct DecoderTC {
    TcConvW frame_in;             // [content 768 | e_fr | log f0] -> [SourceNet 128 | FilterNet 384]
    struct Cnxt {
        const float *w7 = nullptr, *wb = nullptr, *ln_g = nullptr, *ln_b = nullptr, *grn_g = nullptr, *grn_b = nullptr;
        TcConvW c2, c3;
    } mid[3];
    TcConvW heads;                // 128 -> [kernel 961 | 7 pad | amps 15]
    TcConvW dft_cos, dft_sin;     // inverse real-DFT bases, 961 -> 961
    TcConvW down0;
    struct Down { TcConvW c1, c2, c3; } down[4];
    struct Up { TcConvW c1, c2, c3, c4, c5; } up[5];
    const float *out_w = nullptr, *out_b = nullptr;
    float* w7_buf = nullptr;      // depth-wise weights repacked [3][7][128]
    unsigned long long* rng_state = nullptr;   // device {seed, step} of the noise generator (used when no draw is injected)
    bool ready = false;
    ~DecoderTC();
    int init(const WeightStore& store);
    // Decoder.infer (decoder.py:253-257) on the tensor-core path.
    // sink: nullptr = chain tables go into the workspace (eager call); else see ChainSink
    int infer(Arena& A, cudaStream_t s, const float* content, const float* f0, const float* energy, const float* rand01,
              float* out, int B, int Lf, ChainSink* sink = nullptr) const;
};

void set_fused_up(bool on);          // tvc_set_option("fused_up", "0"|"1"): fused 24-channel Upsample block (default on)
bool fused_up();
void set_chain(bool on);             // tvc_set_option("chain", "0"|"1"): low-rate Down / Up blocks as one persistent launch
unsigned plan_options();             // bit set of the options that change the launch plan (part of the graph-cache key)

}  // namespace tvc
