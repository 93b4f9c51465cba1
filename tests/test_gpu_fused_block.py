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


@pytest.mark.parametrize("B,Lf,t0,t1", [(2, 28, 3840, 9600), (3, 7, 0, 1), (3, 7, 3359, 3360), (1, 1, 100, 300), (5, 18, 425, 427),
                                        (4, 28, 0, 13440), (2, 5, 426, 852), (1, 100, 20000, 21000), (2, 60, 0, 3000),
                                        (2, 60, 26000, 28800), (3, 40, 9000, 9100)])
def test_output_pruning_is_bit_identical_inside_the_kept_range(cuda_models, B, Lf, t0, t1):
    """tvc_decoder_infer_range / Decoder.infer(keep=(t0, t1)): the fused block walks only the 426-sample windows that produce
    kept samples (a streaming tick reads y[-9600:-3840] of 13 440, module/infer/stream.py:75).  The walked windows see complete
    inputs, so the kept samples are the bits of the full run -- eager and under graph replay."""
    from tinyvc_b200 import synth
    _, dec = cuda_models
    inp = {k: v.to("cuda") for k, v in synth.decoder_inputs(B, Lf, 77 + Lf + t0).items()}
    full = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"]).clone()
    out = torch.full_like(full, float("nan"))
    for _ in range(3):        # eager, capture, replay
        out.fill_(float("nan"))
        dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"], out=out, keep=(t0, t1))
        assert torch.equal(out[:, t0:t1], full[:, t0:t1])
    # what lies outside the walked windows was not touched (the pruning really skips work)
    lo, hi = (t0 // 426) * 426, min(((t1 - 1) // 426 + 1) * 426, full.shape[1])
    assert torch.isnan(out[:, :lo]).all() and torch.isnan(out[:, hi:]).all()
    with pytest.raises(RuntimeError):
        dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"], keep=(t1, t0))
