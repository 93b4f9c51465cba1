// Fused Upsample block of the FilterNet's highest rate (24 channels): c1 -> c2+FiLM1+residual -> c3 -> c4+FiLM2+residual
// -> c5 in ONE kernel, intermediates in shared memory / TMEM (module/tinyvc/decoder.py:165-171).  EXPERIMENTAL: off
// unless tvc_set_option("fused_up", "1"); see tc_block.cu for its state.
#pragma once
#include "tc_conv.cuh"

namespace tvc {

struct TcUpBlockArgs {
    const bf16 *p_hi = nullptr, *p_lo = nullptr;       // lrelu(x) planes of the up-sampled input, 24 channels of capacity
    const bf16 *c_hi = nullptr, *c_lo = nullptr;       // skip tensor (FiLM condition) planes, 24 channels of capacity
    const float* xi = nullptr;                         // up-sampled input, fp32 chunk-major (residual of c2), 24 channels
    float* xo = nullptr;                               // c5 output, fp32 chunk-major, xo_cs channels of capacity
    int xo_cs = 0;
    int B = 0, T = 0;
    int dil[4] = {1, 3, 9, 27};                        // dilations of c1..c4 (c5 is 1x1)
};

// c1..c5: the block's packed convs (tc_pack_conv; c2 / c4 with their TC_AUX_FILM images).
int tc_up24_block_launch(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5,
                         const TcUpBlockArgs& a, cudaStream_t s);
// whether the packed shapes are the ones the kernel is specialised for
bool tc_up24_block_supported(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5);

}  // namespace tvc
