#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on torch-CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference imports `torchfcpe` and `pyworld` at module scope (module/utils/__init__.py:2 ->
f0_estimation.py:6,9) although the inference path never calls them; both are absent here, so two
empty stub modules are registered before the import.  Nothing from the reference
is copied into this repo -- only the tensors it produces.

Weights are `tinyvc_b200.weights.synth_state_dict(seed)` loaded into the reference's own
Encoder()/Decoder(); inputs are `tinyvc_b200.synth.*`.  Fixtures record torch's version.
"""
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.dont_write_bytecode = True
for _n in ("torchfcpe", "pyworld"):
    _m = types.ModuleType(_n)
    _m.spawn_bundled_infer_model = lambda *a, **k: None
    sys.modules[_n] = _m
# The reference's `module` is a namespace package (no __init__.py), so the repo's own `module/`
# shim would shadow it wherever it sits on sys.path: import the reference FIRST, with the repo
# root not importable, and only then add the repo root for `tinyvc_b200`.
sys.path = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != REPO]
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from module.tinyvc import Encoder, Decoder, match_features  # noqa: E402  (reference)
from module.infer import Generator, StreamInfer  # noqa: E402  (reference)
from module.utils import spectrogram, estimate_energy, shift_frequency, autopad_waveform  # noqa: E402
import module as _ref_module  # noqa: E402

assert list(_ref_module.__path__)[0].startswith(REF), _ref_module.__path__
sys.path.append(REPO)

from tinyvc_b200.weights import synth_state_dict, state_checksum  # noqa: E402
from tinyvc_b200 import synth  # noqa: E402

WEIGHT_SEED = 7
torch.set_num_threads(8)


def build():
    enc, dec = Encoder().eval(), Decoder().eval()
    enc.load_state_dict(synth_state_dict(enc.state_dict(), WEIGHT_SEED))
    dec.load_state_dict(synth_state_dict(dec.state_dict(), WEIGHT_SEED))
    return enc, dec


def meta(enc, dec):
    return dict(torch_version=np.array(torch.__version__), weight_seed=np.array(WEIGHT_SEED),
                enc_checksum=np.array(state_checksum(enc.state_dict())),
                dec_checksum=np.array(state_checksum(dec.state_dict())))


def npy(t):
    return t.detach().cpu().numpy()


@torch.inference_mode()
def golden_decoder(enc, dec):
    """Decoder.infer on B=2, Lf=18 (the bench workload's chunk shape) with a known noise draw."""
    inp = synth.decoder_inputs(2, 18, seed=1235)
    torch.manual_seed(4242)
    rand01 = torch.rand(2, 961, 18)              # what decoder.py:78 will draw under this seed
    amps, kern = dec.source_net(inp["content"], inp["f0"], inp["energy"])
    torch.manual_seed(4242)
    src = dec.dsp(inp["f0"], amps, kern)
    torch.manual_seed(4242)
    out = dec.infer(inp["content"], inp["f0"], inp["energy"])
    np.savez(os.path.join(HERE, "decoder_b2_lf18.npz"), content=npy(inp["content"]), f0=npy(inp["f0"]),
             energy=npy(inp["energy"]), rand01=npy(rand01), amps=npy(amps), kernel=npy(kern),
             source_b0=npy(src[0]), out=npy(out), **meta(enc, dec))
    print("decoder:", out.shape, float(out.pow(2).mean().sqrt()))


@torch.inference_mode()
def golden_pipeline(enc, dec):
    """Generator.convert on ragged-length audio (autopad path) with a 300-vector index."""
    gen = Generator(enc, dec)
    inp = synth.pipeline_inputs(2, 4700, 300, seed=1236)
    wf, index = inp["wf"], inp["index"]
    wfp = autopad_waveform(wf)
    spec = spectrogram(wfp)
    energy = estimate_energy(wfp)
    z, f0 = enc.infer(spec)
    logits = enc.pitch_estimator(spec)
    tgt = index.expand(2, -1, -1)                # bmm needs the batch expanded
    zm = match_features(z, tgt)
    sims = torch.bmm((z.transpose(1, 2) / (torch.norm(z.transpose(1, 2), dim=2, keepdim=True) + 1e-6)),
                     (tgt.transpose(1, 2) / (torch.norm(tgt.transpose(1, 2), dim=2, keepdim=True) + 1e-6)).transpose(1, 2))
    idx = torch.topk(sims, 4, dim=2).indices
    f0s = shift_frequency(f0, 3.0)
    torch.manual_seed(99)
    rand01 = torch.rand(2, 961, spec.shape[2])
    torch.manual_seed(99)
    out = gen.convert(wf, tgt, 3.0)
    tz, tf0 = gen.encode(wf)
    np.savez(os.path.join(HERE, "pipeline_b2_t4700.npz"), wf=npy(wf), index=npy(index), spec=npy(spec),
             energy=npy(energy), z=npy(z), f0=npy(f0), logits_b0=npy(logits[0]), zm=npy(zm), idx=npy(idx),
             f0s=npy(f0s), rand01=npy(rand01), out=npy(out), enc_z=npy(tz), enc_f0=npy(tf0),
             pitch_shift=np.array(3.0), **meta(enc, dec))
    print("pipeline:", out.shape, float(out.pow(2).mean().sqrt()), "f0 range", float(f0.min()), float(f0.max()))


@torch.inference_mode()
def golden_match(enc, dec):
    """match_features: all three metrics and an alpha blend."""
    g = torch.Generator().manual_seed(55)
    src = torch.randn(1, 768, 23, generator=g)
    ref = torch.randn(1, 768, 130, generator=g)
    out = {}
    for m in ("cos", "IP", "L2"):
        out["out_" + m] = npy(match_features(src, ref, k=4, alpha=0.0, metrics=m))
    out["out_cos_a03"] = npy(match_features(src, ref, k=4, alpha=0.3, metrics="cos"))
    out["out_cos_k2"] = npy(match_features(src, ref, k=2, alpha=0.0, metrics="cos"))
    np.savez(os.path.join(HERE, "match_features.npz"), source=npy(src), reference=npy(ref), **out,
             torch_version=np.array(torch.__version__))
    print("match: ok")


@torch.inference_mode()
def golden_stream(enc, dec):
    """StreamInfer.audio_callback over 4 ticks (SOLA cross-fade) and 2 with the phase vocoder."""
    gen = Generator(enc, dec)
    g = torch.Generator().manual_seed(77)
    index = torch.randn(1, 768, 64, generator=g)
    blocks = 0.1 * torch.randn(4, 1920, generator=g)
    res = {}
    for tag, pv in (("sola", False), ("pv", True)):
        si = StreamInfer(gen, target=index, pitch_shift=0.0, use_phase_vocoder=pv)
        si.init_buffer()
        torch.manual_seed(777)
        outs = []
        for i in range(4 if not pv else 2):
            outs.append(si.audio_callback(blocks[i].clone()).clone())
        res["out_" + tag] = npy(torch.stack(outs))
    np.savez(os.path.join(HERE, "stream_4ticks.npz"), index=npy(index), blocks=npy(blocks),
             rand_seed=np.array(777), **res, **meta(enc, dec))
    print("stream:", res["out_sola"].shape, res["out_pv"].shape)


@torch.inference_mode()
def golden_conditioning(enc, dec):
    """How well-conditioned is the reference's own free-running pipeline?  The same Generator.convert call as
    golden_pipeline, once more with 1 CPU thread instead of 8: the BLAS / conv kernels then sum in another order, z and f0
    move by ~1e-6 relative, and the phase integrator of the oscillator (decoder.py:49-50: cumsum of f0 / 24000 over the
    whole utterance, times 15 harmonics) turns that into a visible waveform difference.  Stored: the 1-thread outputs and
    the spread between the two runs, which is the yardstick for any free-running comparison against the reference."""
    gen = Generator(enc, dec)
    inp = synth.pipeline_inputs(2, 4700, 300, seed=1236)
    wf, index = inp["wf"], inp["index"]
    tgt = index.expand(2, -1, -1)
    outs, zs, f0s = {}, {}, {}
    for nt in (8, 1):
        torch.set_num_threads(nt)
        torch.manual_seed(99)
        outs[nt] = gen.convert(wf, tgt, 3.0)
        zs[nt], f0s[nt] = gen.encode(wf)
    torch.set_num_threads(8)
    d = outs[8] - outs[1]
    spread = float(d.pow(2).mean().sqrt())
    z_rel = float((zs[8] - zs[1]).abs().max() / zs[8].abs().max())
    f0_rel = float(((f0s[8] - f0s[1]).abs() / f0s[8].abs().clamp_min(1.0)).max())
    np.savez(os.path.join(HERE, "pipeline_conditioning.npz"), out_8threads=npy(outs[8]), out_1thread=npy(outs[1]),
             f0_8threads=npy(f0s[8]), f0_1thread=npy(f0s[1]), spread_rmse=np.array(spread), z_rel_max=np.array(z_rel),
             f0_rel_max=np.array(f0_rel), **meta(enc, dec))
    print(f"conditioning: reference 8 threads vs 1 thread: waveform RMSE {spread:.3e}, z rel max {z_rel:.2e}, f0 rel max {f0_rel:.2e}, "
          f"f0 range {float(f0s[8].min()):.0f}..{float(f0s[8].max()):.0f} Hz")


@torch.inference_mode()
def golden_resample():
    """torchaudio.functional.resample exactly as reference infer.py:63-64 calls it (defaults), for the rates input files come
    in: 44.1 / 48 / 16 / 22.05 / 8 kHz -> 24 kHz, stereo and mono, lengths that do and do not divide."""
    from torchaudio.functional import resample     # the reference's own import (infer.py:9)
    import torchaudio
    g = torch.Generator().manual_seed(2024)
    out = {"torchaudio_version": np.array(torchaudio.__version__)}
    cases = [(44100, 2, 4410), (48000, 1, 4801), (16000, 2, 1777), (22050, 1, 2205), (8000, 1, 331), (24000, 1, 100), (32000, 1, 7)]
    out["cases"] = np.array(cases)
    for sr, ch, n in cases:
        wf = 0.3 * torch.randn(ch, n, generator=g)
        out[f"in_{sr}"] = npy(wf)
        out[f"out_{sr}"] = npy(resample(wf, sr, 24000))
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)
    print("resample:", {sr: out[f"out_{sr}"].shape for sr, _, _ in cases})


def golden_keys(enc, dec):
    """state_dict key order and shapes of the reference's Encoder / Decoder."""
    import json
    doc = {"encoder": [[k, list(v.shape)] for k, v in enc.state_dict().items()],
           "decoder": [[k, list(v.shape)] for k, v in dec.state_dict().items()]}
    with open(os.path.join(HERE, "state_keys.json"), "w") as f:
        json.dump(doc, f, indent=0)
    print("keys:", len(doc["encoder"]), len(doc["decoder"]))


if __name__ == "__main__":
    enc, dec = build()
    if "--resample-only" in sys.argv:
        golden_resample()
        sys.exit(0)
    if "--conditioning-only" in sys.argv:
        golden_conditioning(enc, dec)
        sys.exit(0)
    golden_conditioning(enc, dec)
    golden_keys(enc, dec)
    golden_decoder(enc, dec)
    golden_pipeline(enc, dec)
    golden_match(enc, dec)
    golden_stream(enc, dec)
    golden_resample()
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("fixtures total bytes:", tot)
