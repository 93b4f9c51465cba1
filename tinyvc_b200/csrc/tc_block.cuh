// Fused Upsample block of the FilterNet's highest rate (24 channels) + output layer: x5 resampler -> c1 -> c2+FiLM1+residual
// -> c3 -> c4+FiLM2+residual -> c5 -> output_layer (k = 7) in ONE kernel, intermediates in shared memory / TMEM
// (module/tinyvc/decoder.py:165-190,220,233).  tvc_set_option("fused_up", "0") runs the separate launches instead.
#pragma once
#include "tc_conv.cuh"

namespace tvc {

struct TcUpBlockArgs {
    const float* x4 = nullptr;                         // block input BEFORE the x5 resampler: fp32 chunk-major, B * T4 rows, 24 channels
    const bf16 *c_hi = nullptr, *c_lo = nullptr;       // skip tensor (FiLM condition) planes, B * T rows, 24 channels of capacity
    const float *out_w = nullptr, *out_b = nullptr;    // output_layer.weight [1][24][7] / .bias [1] (torch layout)
    float* out = nullptr;                              // waveform [B][T]
    int B = 0, T = 0, T4 = 0;                          // T = 5 * T4
    float scale = 0.2f;                                // F.interpolate's source-index scale, (float)(1 / 5)
    int dil[4] = {1, 3, 9, 27};                        // dilations of c1..c4 (c5 is 1x1)
    // Output pruning: only out[b][t_lo, t_hi) has to be produced (t_hi < 0: the whole utterance).  Windows that produce none
    // of those samples are not walked; the windows that are walked compute exactly what they compute in a full run (their
    // inputs x4 / cond are complete tensors), so the samples written are bit-identical to the full run's.  A streaming tick
    // keeps 5 760 of its 13 440 samples (module/infer/stream.py:75).
    int t_lo = 0, t_hi = -1;
    // x4 may hold only rows [x4_off, x4_off + x4_rows) of every utterance's low-rate tensor (x4_rows <= 0: all T4 rows): the
    // rows the walked windows resample must lie inside (the plan computes the range, nets_tc.cu)
    int x4_rows = 0, x4_off = 0;
};

// c1..c5: the block's packed convs (tc_pack_conv; c2 / c4 with their TC_AUX_FILM images).
int tc_up24_block_launch(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5,
                         const TcUpBlockArgs& a, cudaStream_t s);
// whether the packed shapes are the ones the kernel is specialised for
bool tc_up24_block_supported(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5);

}  // namespace tvc
