"""Fused 24-channel Upsample block (csrc/tc_block.cu, module/tinyvc/decoder.py:165-190) against the five tc_conv launches.

The fused kernel issues the same MMAs per tile in the same order and runs the same epilogue arithmetic, so the waveform
must be bit-identical to the unfused plan (tvc_set_option("fused_up", "0")) for every shape: windows of 432 output rows
with 40-row halos, first / last windows clamped to the utterance (replicate padding), utterances shorter than a window.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,Lf", [(1, 1), (2, 1), (3, 2), (5, 7), (64, 18), (2, 100)])
def test_fused_block_bit_identical(cuda_models, B, Lf):
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
    assert torch.equal(got, ref), f"max|d| = {float((got - ref).abs().max()):.3e}"
