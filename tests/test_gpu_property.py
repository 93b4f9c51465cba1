"""Property test of Decoder.infer over the shape / voicing space (hypothesis): any frame count 1..600, any batch 1..3, f0
contours with exact zeros, sub-threshold values (<= 20 Hz counts as unvoiced, decoder.py:45) and arbitrary unvoiced runs,
including all-unvoiced and all-voiced utterances.  Properties checked against the CPU oracle (oracle/tinyvc_oracle.py, the
pinned restatement of module/tinyvc/decoder.py:24-85,193-266) with the same injected noise draw:
waveform RMSE < 1e-4 (the north star's bar), finite everywhere, and independence of the batch."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from conftest import rmse
from oracle import tinyvc_oracle as O
from tinyvc_b200 import synth

pytestmark = pytest.mark.gpu


@st.composite
def decoder_case(draw):
    lf = draw(st.one_of(st.integers(1, 8), st.integers(9, 120), st.integers(121, 600)))
    batch = draw(st.integers(1, 3))
    seed = draw(st.integers(0, 2**20))
    mode = draw(st.sampled_from(["mixed", "all_unvoiced", "all_voiced", "sub_threshold", "runs"]))
    return lf, batch, seed, mode


def _f0(inp, lf, batch, seed, mode):
    g = torch.Generator().manual_seed(seed + 17)
    f0 = inp["f0"].clone()
    if mode == "all_unvoiced":
        f0.zero_()
    elif mode == "all_voiced":
        f0 = f0.clamp_min(80.0)
    elif mode == "sub_threshold":                       # values in (0, 20] are unvoiced for the mask but still drive the phase
        m = torch.rand(batch, 1, lf, generator=g) < 0.4
        f0 = torch.where(m, 20.0 * torch.rand(batch, 1, lf, generator=g), f0)
    elif mode == "runs":                                # unvoiced runs of random length at random places, exact zeros
        for b in range(batch):
            pos = 0
            while pos < lf:
                n = int(torch.randint(1, 12, (1,), generator=g))
                if float(torch.rand(1, generator=g)) < 0.5:
                    f0[b, 0, pos:pos + n] = 0.0
                pos += n
    return f0


@settings(max_examples=12, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.data_too_large])
@given(case=decoder_case())
def test_decoder_infer_matches_oracle_for_any_shape_and_voicing(cuda_models, weights, case):
    lf, batch, seed, mode = case
    _, dec = cuda_models
    PD = weights[1]
    inp = synth.decoder_inputs(batch, lf, seed=seed)
    f0 = _f0(inp, lf, batch, seed, mode)
    out = dec.infer(inp["content"].cuda(), f0.cuda(), inp["energy"].cuda(), rand01=inp["rand01"].cuda())
    assert out.shape == (batch, lf * 480) and bool(torch.isfinite(out).all())
    ref = O.decoder_infer(PD, inp["content"], f0, inp["energy"], inp["rand01"])
    e = rmse(out, ref)
    assert e < 1e-4, (case, e)
    if batch > 1:                                       # an utterance does not see its neighbours in the batch
        b = batch - 1
        alone = dec.infer(inp["content"][b:b + 1].cuda(), f0[b:b + 1].cuda(), inp["energy"][b:b + 1].cuda(),
                          rand01=inp["rand01"][b:b + 1].cuda())
        assert torch.equal(alone[0], out[b]), case


@st.composite
def keep_case(draw):
    lf = draw(st.one_of(st.integers(1, 12), st.integers(13, 60), st.integers(61, 160)))
    L = lf * 480
    a = draw(st.integers(0, L - 1))
    b = draw(st.integers(a + 1, L))
    # bias towards short ranges in long utterances (where the levels below the block get pruned) and towards the two ends
    kind = draw(st.sampled_from(["any", "short", "head", "tail"]))
    if kind == "short":
        b = min(L, a + draw(st.integers(1, 4000)))
    elif kind == "head":
        a = 0
    elif kind == "tail":
        b = L
    return lf, draw(st.integers(1, 3)), draw(st.integers(0, 2**20)), a, b


@settings(max_examples=25, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.data_too_large])
@given(case=keep_case())
def test_output_pruning_is_exact_for_any_range(cuda_models, case):
    """Decoder.infer(keep=(t0, t1)) for arbitrary utterance lengths and ranges: the kept samples are the full run's bits (the
    fused block walks only the windows that produce them; on long utterances the Upsample levels below it run on compact
    tensors, nets_tc.cu)."""
    lf, batch, seed, t0, t1 = case
    _, dec = cuda_models
    inp = {k: v.cuda() for k, v in synth.decoder_inputs(batch, lf, seed=seed).items()}
    full = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"]).clone()
    part = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"], keep=(t0, t1))
    assert torch.equal(part[:, t0:t1], full[:, t0:t1]), (case, float((part[:, t0:t1] - full[:, t0:t1]).abs().max()))
