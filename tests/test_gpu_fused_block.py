"""Fused full-rate block (csrc/tc_block.cu: x5 resampler + 24-channel Upsample block + output layer,
module/tinyvc/decoder.py:165-190,220,233) against the separate launches (tvc_set_option("fused_up", "0")).

Same resampler formula, same epilogue arithmetic, same output-conv summation order; the one difference is that the fused
kernel forms x_hi*w_hi + x_lo*w_hi and x_hi*w_lo in two accumulators and adds them once (two MMAs per K-step instead of
three), so the two plans agree to a few fp32 ulps of every intermediate, not bit for bit: max |d| < 2e-5 on waveforms of
RMS ~0.27 (an indexing error -- window halos, replicate rows, utterance edges, short utterances -- shows up as O(0.1)).
The fused plan itself must be reproducible bit for bit (eager run vs graph replay)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,Lf", [(1, 1), (2, 1), (3, 2), (5, 7), (64, 18), (2, 100)])
def test_fused_block_matches_separate_launches(cuda_models, B, Lf):
    from tinyvc_b200 import _lib, synth
    _, dec = cuda_models
    inp = {k: v.to("cuda") for k, v in synth.decoder_inputs(B, Lf, 1236 + B).items()}
    run = lambda: dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"]).clone()
    try:
        _lib.set_option("fused_up", "0")
        ref = run()
        ref2 = run()          # second sighting of the buffer set: captured graph
        _lib.set_option("fused_up", "1")
        got = run()
        got2 = run()
    finally:
        _lib.set_option("fused_up", "1")
    assert torch.equal(ref, ref2)
    assert torch.equal(got, got2)
    assert torch.isfinite(got).all()
    d = float((got - ref).abs().max())
    assert d < 2e-5, f"max|d| = {d:.3e}"


@pytest.mark.parametrize("B,Lf", [(1, 1), (3, 2), (5, 7), (64, 18), (2, 100), (7, 33)])
def test_fused_down_resamplers_bit_identical(cuda_models, B, Lf):
    """The 1/f resamplers in front of the Downsample blocks (decoder.py:148) evaluated inside the producing conv's epilogue
    (tc_conv.cu SPEC 9: factors 5 / 4 / 4 / 3, neighbour row by warp shuffle, interp_cl's coordinate arithmetic) against the
    separate interp_cl launches: identical waveforms, eager and graph replay alike."""
    from tinyvc_b200 import _lib, synth
    _, dec = cuda_models
    inp = {k: v.to("cuda") for k, v in synth.decoder_inputs(B, Lf, 501 + Lf).items()}
    run = lambda: dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"]).clone()
    try:
        _lib.set_option("fuse_down", "0")
        ref = run()
        _lib.set_option("fuse_down", "1")
        got = [run() for _ in range(3)]
    finally:
        _lib.set_option("fuse_down", "1")
    for g in got:
        assert torch.isfinite(g).all()
        assert torch.equal(g, ref), f"max|d| = {float((g - ref).abs().max()):.3e}"
