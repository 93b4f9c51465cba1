"""Decoder parity on the GPU: every stage of Decoder.infer through the C-ABI vs the CPU oracle.

Bar (BASELINE.json north_star): waveform fp32 RMSE < 1e-4 against the reference's CPU path on the
same inputs (same weights, same noise draw); the oscillator phase bit-equal."""
import pytest
import torch

from conftest import load_golden, t, rmse, max_abs
from oracle import tinyvc_oracle as O
from tinyvc_b200 import synth

pytestmark = pytest.mark.gpu
RMSE_BAR = 1e-4


def _cuda(*xs):
    return [x.to("cuda") for x in xs]


@pytest.mark.parametrize("lf,seed", [(18, 0), (57, 1), (200, 2)])
def test_oscillator_phase_bit_exact(lf, seed, report):
    from tinyvc_b200.tinyvc.decoder import harmonic_theta
    f0 = synth.synth_f0(2, lf, torch.Generator().manual_seed(seed))
    _, theta_ref = O.oscillate_harmonics(f0, return_theta=True)
    theta = harmonic_theta(f0.cuda()).cpu()
    nbad = int((theta != theta_ref).sum())
    report.add(f"theta_lf{lf}", mismatches=nbad, max_abs=max_abs(theta, theta_ref))
    assert nbad == 0, f"{nbad} of {theta.numel()} phase samples differ from the CPU reference"


def test_oscillator_phase_high_pitch_bit_exact(report):
    from tinyvc_b200.tinyvc.decoder import harmonic_theta
    f0 = torch.rand(1, 1, 40, generator=torch.Generator().manual_seed(5)) * 14000.0
    _, theta_ref = O.oscillate_harmonics(f0, return_theta=True)
    theta = harmonic_theta(f0.cuda()).cpu()
    nbad = int((theta != theta_ref).sum())
    report.add("theta_high_pitch", mismatches=nbad)
    assert nbad == 0


@torch.inference_mode()
def test_source_net(cuda_models, weights, report):
    _, dec = cuda_models
    g = load_golden("decoder_b2_lf18.npz")
    content, f0, energy = t(g["content"]), t(g["f0"]), t(g["energy"])
    amps, kern = dec.source_net(*_cuda(content, f0, energy))
    ea, ek = max_abs(amps, t(g["amps"])), max_abs(kern, t(g["kernel"]))
    report.add("source_net", amps_max_abs=ea, kernel_max_abs=ek)
    assert ea < 2e-5 and ek < 2e-5


@torch.inference_mode()
def test_dsp(cuda_models, report):
    _, dec = cuda_models
    g = load_golden("decoder_b2_lf18.npz")
    f0, amps, kern, rand01 = t(g["f0"]), t(g["amps"]), t(g["kernel"]), t(g["rand01"])
    src = dec.dsp(*_cuda(f0, amps, kern), rand01=rand01.cuda()).cpu()
    ref = t(g["source_b0"])
    eh, en = max_abs(src[0, :15], ref[:15]), max_abs(src[0, 15], ref[15])
    report.add("dsp", harmonics_max_abs=eh, noise_max_abs=en, noise_rms=float(ref[15].pow(2).mean().sqrt()))
    assert eh < 5e-6, "harmonics"
    assert en < 2e-5, "noise"


@torch.inference_mode()
def test_filter_net(cuda_models, weights, report):
    _, dec = cuda_models
    g = load_golden("decoder_b2_lf18.npz")
    content, f0, energy, rand01 = t(g["content"]), t(g["f0"]), t(g["energy"]), t(g["rand01"])
    src = O.decoder_dsp(f0, t(g["amps"]), t(g["kernel"]), rand01)
    ref = O.filter_net(weights[1], content, f0, energy, src)
    out = dec.filter_net(*_cuda(content, f0, energy, src)).cpu()
    e = rmse(out, ref)
    report.add("filter_net", rmse=e, ref_rms=float(ref.pow(2).mean().sqrt()), max_abs=max_abs(out, ref))
    assert e < RMSE_BAR


@torch.inference_mode()
def test_decoder_infer_golden(cuda_models, report):
    _, dec = cuda_models
    g = load_golden("decoder_b2_lf18.npz")
    out = dec.infer(*_cuda(t(g["content"]), t(g["f0"]), t(g["energy"])), rand01=t(g["rand01"]).cuda()).cpu()
    ref = t(g["out"])
    e = rmse(out, ref)
    report.add("decoder_golden", rmse=e, ref_rms=float(ref.pow(2).mean().sqrt()), max_abs=max_abs(out, ref))
    assert out.shape == ref.shape
    assert e < RMSE_BAR


@torch.inference_mode()
def test_decoder_infer_golden_fp32_plan(cuda_models, report):
    """The exact-fp32 CUDA-core plan (option conv_impl=fp32) stays available and matches to ~1e-7."""
    from tinyvc_b200 import _lib
    _, dec = cuda_models
    g = load_golden("decoder_b2_lf18.npz")
    _lib.set_option("conv_impl", "fp32")
    try:
        out = dec.infer(*_cuda(t(g["content"]), t(g["f0"]), t(g["energy"])), rand01=t(g["rand01"]).cuda()).cpu()
        # without an injected draw the public call draws torch.rand on the device (the fp32 plan has no in-kernel generator)
        free = dec.infer(*_cuda(t(g["content"]), t(g["f0"]), t(g["energy"])))
        assert free.shape == out.shape and torch.isfinite(free).all()
    finally:
        _lib.set_option("conv_impl", "tc")
    e = rmse(out, t(g["out"]))
    report.add("decoder_golden_fp32_plan", rmse=e)
    assert e < 1e-6


@pytest.mark.parametrize("batch,lf", [(3, 1), (2, 2), (2, 3), (2, 5), (3, 33), (1, 100)])
@torch.inference_mode()
def test_decoder_ragged_lengths(cuda_models, weights, batch, lf, report):
    """Edge cases: utterances shorter than the dilations (replicate padding dominates), odd frame counts."""
    _, dec = cuda_models
    inp = synth.decoder_inputs(batch, lf, seed=100 + lf)
    ref = O.decoder_infer(weights[1], inp["content"], inp["f0"], inp["energy"], inp["rand01"])
    out = dec.infer(*_cuda(inp["content"], inp["f0"], inp["energy"]), rand01=inp["rand01"].cuda()).cpu()
    e = rmse(out, ref)
    report.add(f"decoder_b{batch}_lf{lf}", rmse=e, ref_rms=float(ref.pow(2).mean().sqrt()))
    assert e < RMSE_BAR


@torch.inference_mode()
def test_decoder_unvoiced_and_silent(cuda_models, weights, report):
    """All-unvoiced f0 (oscillator gated off) and zero energy."""
    _, dec = cuda_models
    inp = synth.decoder_inputs(2, 12, seed=77)
    inp["f0"].zero_()
    inp["energy"][1].zero_()
    ref = O.decoder_infer(weights[1], inp["content"], inp["f0"], inp["energy"], inp["rand01"])
    out = dec.infer(*_cuda(inp["content"], inp["f0"], inp["energy"]), rand01=inp["rand01"].cuda()).cpu()
    e = rmse(out, ref)
    report.add("decoder_unvoiced", rmse=e)
    assert e < RMSE_BAR


@torch.inference_mode()
def test_decoder_batch_invariance(cuda_models):
    """An utterance's waveform must not depend on what else is in the batch (sharding invariant)."""
    _, dec = cuda_models
    inp = synth.decoder_inputs(5, 18, seed=9)
    c, f, e, r = _cuda(inp["content"], inp["f0"], inp["energy"], inp["rand01"])
    full = dec.infer(c, f, e, rand01=r)
    for b in (0, 3, 4):
        one = dec.infer(c[b:b + 1], f[b:b + 1], e[b:b + 1], rand01=r[b:b + 1])
        assert torch.equal(one[0], full[b]), f"utterance {b} differs when run alone"
    again = dec.infer(c, f, e, rand01=r)
    assert torch.equal(again, full), "run-to-run nondeterminism"


@torch.inference_mode()
def test_bench_shape_properties(cuda_models, weights, report):
    """BASELINE config 2 shape (B=64, Lf=18): oracle parity on 8 utterances, finite everywhere."""
    _, dec = cuda_models
    inp = synth.decoder_inputs(64, 18, seed=1234 + 2)
    out = dec.infer(*_cuda(inp["content"], inp["f0"], inp["energy"]), rand01=inp["rand01"].cuda()).cpu()
    assert out.shape == (64, 8640) and torch.isfinite(out).all()
    sel = [0, 9, 18, 27, 36, 45, 54, 63]
    ref = O.decoder_infer(weights[1], inp["content"][sel], inp["f0"][sel], inp["energy"][sel], inp["rand01"][sel])
    e = rmse(out[sel], ref)
    report.add("decoder_config2_8utt", rmse=e, ref_rms=float(ref.pow(2).mean().sqrt()))
    assert e < RMSE_BAR


def test_cpu_tensor_is_rejected(cuda_models):
    _, dec = cuda_models
    inp = synth.decoder_inputs(1, 4, seed=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        dec.infer(inp["content"], inp["f0"], inp["energy"])


@torch.inference_mode()
def test_decoder_graph_replay_matches_eager(cuda_models):
    """tvc_decoder_infer captures a CUDA graph the second time it sees a buffer set; replays must read the
    buffers' *current* contents and reproduce the eager launch sequence bit for bit."""
    from tinyvc_b200 import _lib, synth
    _, dec = cuda_models
    a = {k: v.cuda() for k, v in synth.decoder_inputs(3, 7, seed=21).items()}
    b = {k: v.cuda() for k, v in synth.decoder_inputs(3, 7, seed=22).items()}
    _lib.set_option("graphs", "0")
    try:
        want_a = dec.infer(a["content"], a["f0"], a["energy"], rand01=a["rand01"]).clone()
        want_b = dec.infer(b["content"], b["f0"], b["energy"], rand01=b["rand01"]).clone()
    finally:
        _lib.set_option("graphs", "1")
    L, h = _lib.lib(), dec._native.get()
    out = torch.empty_like(want_a)
    ws = torch.empty(L.tvc_decoder_workspace_bytes(3, 7), dtype=torch.uint8, device="cuda")
    stat = {k: v.clone() for k, v in a.items()}
    n0 = _lib.launch_count()
    per_call = []
    for it in range(4):                       # call 0 eager, call 1 captures, calls 2.. replay
        src = a if it % 2 == 0 else b
        for k in stat:
            stat[k].copy_(src[k])
        _lib.check(L.tvc_decoder_infer(h, stat["content"].data_ptr(), stat["f0"].data_ptr(), stat["energy"].data_ptr(),
                                       stat["rand01"].data_ptr(), out.data_ptr(), 3, 7, ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr(out.device)), "tvc_decoder_infer")
        torch.cuda.synchronize()
        assert torch.equal(out, want_a if it % 2 == 0 else want_b), f"call {it} differs from the eager result"
        per_call.append(_lib.launch_count() - n0)
        n0 = _lib.launch_count()
    assert len(set(per_call)) == 1 and per_call[0] > 50, per_call   # replays account for every captured kernel


@torch.inference_mode()
def test_in_kernel_noise_draw(cuda_models, weights, report):
    """rand01 omitted: the uniform draw of decoder.py:78 is made inside the noise kernel (Philox, seeded per decoder).
    Same seed -> same waveform; consecutive calls -> fresh draws (also under CUDA-graph replay); the result differs
    from an oracle run by what two oracle runs with different draws differ by (same noise power)."""
    _, dec = cuda_models
    PD = weights[1]
    inp = synth.decoder_inputs(3, 12, seed=41)
    c, f, e = inp["content"].cuda(), inp["f0"].cuda(), inp["energy"].cuda()
    dec.seed_noise(1234)
    a1 = dec.infer(c, f, e).clone()
    a2 = dec.infer(c, f, e).clone()
    a3 = dec.infer(c, f, e).clone()          # third call with the same buffers: graph replay
    dec.seed_noise(1234)
    b1 = dec.infer(c, f, e).clone()
    b2 = dec.infer(c, f, e).clone()
    assert torch.equal(a1, b1) and torch.equal(a2, b2), "same seed must reproduce the same sequence of draws"
    assert not torch.equal(a1, a2) and not torch.equal(a2, a3), "every call must use a fresh draw"
    g = torch.Generator().manual_seed(3)
    ra, rb = torch.rand(3, 961, 12, generator=g), torch.rand(3, 961, 12, generator=g)
    oa = O.decoder_infer(PD, inp["content"], inp["f0"], inp["energy"], ra)
    ob = O.decoder_infer(PD, inp["content"], inp["f0"], inp["energy"], rb)
    between = rmse(oa, ob)
    ours = rmse(a1, oa)
    report.add("in_kernel_noise", rmse_vs_oracle_draw=ours, rmse_between_oracle_draws=between)
    assert 0.5 * between < ours < 2.0 * between


@torch.inference_mode()
def test_infer_into_pinned_host_memory(cuda_models):
    """Decoder.infer(out=pinned host tensor): the last kernel writes the waveform through the PCIe mapping; same bits as the
    device result."""
    _, dec = cuda_models
    inp = {k: v.cuda() for k, v in synth.decoder_inputs(3, 7, seed=31).items()}
    want = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
    host = torch.empty(3, 7 * 480).pin_memory()
    got = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"], out=host)
    torch.cuda.synchronize()
    assert got.data_ptr() == host.data_ptr() and torch.equal(host, want.cpu())
    with pytest.raises(RuntimeError):
        dec.infer(inp["content"], inp["f0"], inp["energy"], out=torch.empty(3, 7 * 480))      # pageable host memory
