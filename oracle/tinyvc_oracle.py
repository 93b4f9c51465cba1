"""CPU oracle for the TinyVC real-time inference path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional (state-dict in, tensors out) restatement, on torch-CPU fp32, of the
algorithm the reference implements in `module/tinyvc`, `module/utils` and `module/infer`.
It exists so the CUDA product path can be checked against "what the reference's -d cpu path
computes" on a machine where `/root/reference` is not present (the GPU box).

Rules (see DESIGN.md "Oracle"):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import this module.  Nothing under `tinyvc_b200/` imports it; the product path has
    no CPU fallback and raises if the CUDA library is missing.
  * the arithmetic of the reference lives in third-party `torch` (un-pinned in the reference's
    requirements.txt:1).  The oracle therefore calls the same ATen CPU ops in the same order, so
    it is bit-identical to the reference on the same torch build.
  * PINNED: `tests/golden/*.npz` hold outputs of the real reference (imported from
    /root/reference by `tests/golden/make_golden.py`, torch 2.11.0+cu128 CPU/AVX512);
    `tests/test_oracle_golden.py` checks this file against them bit-for-bit.  The reference
    itself ships no tests or golden vectors, so that is the only pin.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Parameters are passed as a flat mapping `P` (a state_dict) plus a key prefix.
"""
from __future__ import annotations

import math
from typing import Mapping, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Mapping[str, Tensor]

SAMPLE_RATE = 24000
FRAME = 480
N_FFT = 1920
FFT_BIN = N_FFT // 2 + 1
NUM_HARMONICS = 14          # -> 15 oscillator channels (fundamental + 14)


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def _conv(P: Params, name: str, x: Tensor, *, dilation: int = 1, replicate_pad: int = 0,
          groups: int = 1) -> Tensor:
    """nn.Conv1d forward.  `padding_mode='replicate'` convs pad explicitly then run an un-padded
    conv, which is what torch's `_conv_forward` does for non-zero padding modes."""
    w = P[name + ".weight"]
    b = P[name + ".bias"]
    if replicate_pad:
        x = F.pad(x, (replicate_pad, replicate_pad), mode="replicate")
    return F.conv1d(x, w, b, stride=1, padding=0, dilation=dilation, groups=groups)


def _log_f0(f0: Tensor) -> Tensor:
    # decoder.py:128 and :223  log(relu(f0) + 1e-6)
    return torch.log(F.relu(f0) + 1e-6)


# --------------------------------------------------------------------------------------------
# module/utils  (inference-time signal helpers)
# --------------------------------------------------------------------------------------------
def autopad_waveform(wf: Tensor, frame_size: int = FRAME) -> Tensor:
    """utils/auto_padding.py:5-11 -- right zero-pad [B,T] to a multiple of frame_size."""
    rem = wf.shape[1] % frame_size
    if rem == 0:
        return wf
    return torch.cat([wf, wf.new_zeros(wf.shape[0], frame_size - rem)], dim=1)


def spectrogram(wave: Tensor, n_fft: int = N_FFT, hop_size: int = FRAME) -> Tensor:
    """utils/spectrogram.py:8-15 -- centred (reflect) Hann STFT magnitude, frame 0 dropped."""
    win = torch.hann_window(n_fft)
    s = torch.stft(wave.float(), n_fft, hop_size, window=win, return_complex=True).abs()
    return s[:, :, 1:]


def estimate_energy(wave: Tensor, frame_size: int = 64) -> Tensor:
    """utils/energy_estimation.py:9-14 -- max-pool(|x|,128,64,pad 32) then linear interp to T."""
    e = F.max_pool1d(wave.abs().unsqueeze(1), frame_size * 2, frame_size, frame_size // 2)
    return F.interpolate(e, wave.shape[1], mode="linear")


def shift_frequency(f0: Tensor, shift: float) -> Tensor:
    """utils/pitch_shift.py:5-15 -- semitone shift through a log2 'midi' domain."""
    midi = torch.log2(F.relu(f0 / 440) + 1e-6) * 12 + 69
    midi = midi + shift
    return 440 * 2 ** ((midi - 69) / 12)


# --------------------------------------------------------------------------------------------
# module/tinyvc/convnext.py
# --------------------------------------------------------------------------------------------
def channel_layer_norm(P: Params, name: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """convnext.py:7-19 -- LayerNorm over the channel axis of [B,C,T]."""
    c = x.shape[1]
    y = F.layer_norm(x.transpose(1, 2), (c,), P[name + ".gamma"], P[name + ".beta"], eps)
    return y.transpose(1, 2)


def grn(P: Params, name: str, x: Tensor, eps: float = 1e-6) -> Tensor:
    """convnext.py:23-34 -- global response norm: L2 over time, divided by its channel mean."""
    gx = torch.norm(x, p=2, dim=2, keepdim=True)
    nx = gx / (gx.mean(dim=1, keepdim=True) + eps)
    return P[name + ".gamma"] * (x * nx) + P[name + ".beta"] + x


def convnext_layer(P: Params, name: str, x: Tensor, kernel_size: int = 7, dilation: int = 1) -> Tensor:
    """convnext.py:38-58 -- dw-conv(k, dil, replicate) -> LN -> 1x1 (x2) -> GELU -> GRN -> 1x1 -> +res."""
    c = x.shape[1]
    pad = (kernel_size * dilation - dilation) // 2
    y = _conv(P, name + ".c1", x, dilation=dilation, replicate_pad=pad, groups=c)
    y = channel_layer_norm(P, name + ".norm", y)
    y = _conv(P, name + ".c2", y)
    y = F.gelu(y)
    y = grn(P, name + ".grn", y)
    y = _conv(P, name + ".c3", y)
    return y + x


# --------------------------------------------------------------------------------------------
# module/tinyvc/encoder.py
# --------------------------------------------------------------------------------------------
SSL_DILATIONS = (1, 3, 9, 1, 1, 1)      # encoder.py:80
PITCH_LAYERS = 4                        # encoder.py:16
PITCH_CLASSES = 512
PITCH_CPO = 48
PITCH_FMIN = 20.0


def ssl_features(P: Params, spec: Tensor, prefix: str = "ssl_feature_estimator") -> Tensor:
    """encoder.py:89-97 -- 961->384, LN, 6 ConvNeXt(384), 384->768."""
    x = _conv(P, prefix + ".input_layer", spec)
    x = channel_layer_norm(P, prefix + ".norm", x)
    for i, d in enumerate(SSL_DILATIONS):
        x = convnext_layer(P, f"{prefix}.mid_layers.{i}", x, dilation=d)
    return _conv(P, prefix + ".output_layer", x)


def pitch_logits(P: Params, spec: Tensor, prefix: str = "pitch_estimator") -> Tensor:
    """encoder.py:33-38 -- 961->128, LN, 4 ConvNeXt(128), 128->512 logits."""
    x = _conv(P, prefix + ".input_layer", spec)
    x = channel_layer_norm(P, prefix + ".norm", x)
    for i in range(PITCH_LAYERS):
        x = convnext_layer(P, f"{prefix}.mid_layers.{i}", x)
    return _conv(P, prefix + ".output_layer", x)


def id2freq(ids: Tensor) -> Tensor:
    """encoder.py:48-54 -- 20 * 2^(id/48), values <= 20 Hz forced to 0."""
    x = ids.to(torch.float)
    x = PITCH_FMIN * (2 ** (x / PITCH_CPO))
    x[x <= PITCH_FMIN] = 0
    return x


def pitch_decode(logits: Tensor, k: int = 4) -> Tensor:
    """encoder.py:61-67 -- softmax over the top-k logits, expectation of their frequencies."""
    top, idx = torch.topk(logits, k, dim=1)
    p = F.softmax(top, dim=1)
    f0 = (p * id2freq(idx)).sum(dim=1, keepdim=True)
    f0[f0 <= PITCH_FMIN] = 0
    return f0


def encoder_infer(P: Params, spec: Tensor) -> Tuple[Tensor, Tensor]:
    """encoder.py:113-116 -- (content z [B,768,Lf], f0 [B,1,Lf])."""
    return ssl_features(P, spec), pitch_decode(pitch_logits(P, spec))


# --------------------------------------------------------------------------------------------
# module/tinyvc/feature_retrieval.py
# --------------------------------------------------------------------------------------------
def match_features(source: Tensor, reference: Tensor, k: int = 4, alpha: float = 0.0,
                   metrics: str = "cos", return_indices: bool = False):
    """feature_retrieval.py:15-33.  Runs one utterance at a time (the reference's bmm needs the
    index expanded to the batch and then materialises [B,N,768] + [B,Lf,N]; per-utterance
    execution is bit-identical, SURVEY.md 8c 'Oracle batching')."""
    outs, idxs = [], []
    ref_t = reference[0].transpose(0, 1)                         # [N, C]
    for b in range(source.shape[0]):
        s = source[b].transpose(0, 1)                            # [Lf, C]
        r = ref_t if reference.shape[0] == 1 else reference[b].transpose(0, 1)
        if metrics == "IP":
            sims = torch.bmm(s[None], r.transpose(0, 1)[None])[0]
        elif metrics == "L2":
            sims = -torch.cdist(s[None], r[None])[0]
        elif metrics == "cos":
            rn = torch.norm(r[None], dim=2, keepdim=True, p=2) + 1e-6
            sn = torch.norm(s[None], dim=2, keepdim=True, p=2) + 1e-6
            sims = torch.bmm(s[None] / sn, (r[None] / rn).transpose(1, 2))[0]
        else:
            raise ValueError(metrics)
        best = torch.topk(sims, k, dim=1)
        outs.append(r[best.indices].mean(dim=1).transpose(0, 1))  # [C, Lf]
        idxs.append(best.indices)
    res = torch.stack(outs, dim=0)
    res = res * (1 - alpha) + source * alpha
    if return_indices:
        return res, torch.stack(idxs, dim=0)
    return res


# --------------------------------------------------------------------------------------------
# module/tinyvc/decoder.py -- DSP source
# --------------------------------------------------------------------------------------------
def oscillate_harmonics(f0: Tensor, frame_size: int = FRAME, sample_rate: int = SAMPLE_RATE,
                        num_harmonics: int = NUM_HARMONICS, min_frequency: float = 20.0,
                        return_theta: bool = False):
    """decoder.py:24-54.  NB torch.cumsum on CPU fp32 accumulates in fp64 and rounds every
    prefix to fp32; the `% 1` then sees that rounded value."""
    lw = f0.shape[2] * frame_size
    mul = (torch.arange(num_harmonics + 1) + 1).unsqueeze(0).unsqueeze(2)
    fs = F.interpolate(f0, lw, mode="linear") * mul
    uv = F.interpolate((f0 > min_frequency).to(torch.float), lw, mode="linear")
    integ = torch.cumsum(fs / sample_rate, dim=2)
    theta = 2 * math.pi * (integ % 1)
    h = torch.sin(theta) * uv
    return (h, theta) if return_theta else h


def noise_angle(rand01: Tensor) -> Tensor:
    """decoder.py:78 -- `torch.rand(...) * 2 * math.pi - math.pi` with the uniform draw injected."""
    return rand01 * 2 * math.pi - math.pi


def oscillate_noise(kernel: Tensor, rand01: Tensor, frame_size: int = FRAME, n_fft: int = N_FFT) -> Tensor:
    """decoder.py:63-85 with the `torch.rand` draw passed in as `rand01` [B,961,Lf]."""
    kernel = kernel.to(torch.float)
    y = torch.exp(1j * noise_angle(rand01)) * kernel
    y = F.pad(y, [1, 0])
    out = torch.istft(y, n_fft, frame_size, window=torch.ones(n_fft))
    return out.unsqueeze(1)


def decoder_dsp(f0: Tensor, amps: Tensor, kernel: Tensor, rand01: Tensor) -> Tensor:
    """decoder.py:259-266 -- [B,16,L] = cat(harmonics * interp(amps), noise)."""
    h = oscillate_harmonics(f0)
    a = F.interpolate(amps, scale_factor=FRAME, mode="linear")
    return torch.cat([h * a, oscillate_noise(kernel, rand01)], dim=1)


# --------------------------------------------------------------------------------------------
# module/tinyvc/decoder.py -- networks
# --------------------------------------------------------------------------------------------
SOURCE_LAYERS = 3
FILTER_CHANNELS = (384, 192, 96, 48, 24)     # decoder.py:195
FILTER_FACTORS = (2, 3, 4, 4, 5)             # decoder.py:196


def source_net(P: Params, content: Tensor, f0: Tensor, energy: Tensor,
               prefix: str = "source_net") -> Tuple[Tensor, Tensor]:
    """decoder.py:126-134 -- amps [B,15,Lf], kernel [B,961,Lf]."""
    e = F.max_pool1d(energy, FRAME, FRAME)
    x = _conv(P, prefix + ".content_in", content) + _conv(P, prefix + ".energy_in", e) \
        + _conv(P, prefix + ".f0_in", _log_f0(f0))
    for i in range(SOURCE_LAYERS):
        x = convnext_layer(P, f"{prefix}.mid_layers.{i}", x)
    amps = F.elu(_conv(P, prefix + ".to_amps", x)) + 1.0
    kern = F.elu(_conv(P, prefix + ".to_kernel", x)) + 1.0
    return amps, kern


def film(P: Params, name: str, x: Tensor, c: Tensor) -> Tensor:
    """decoder.py:88-97."""
    shift = _conv(P, name + ".to_shift", c)
    scale = _conv(P, name + ".to_scale", c)
    return x * scale + shift


def downsample_block(P: Params, name: str, x: Tensor, factor: int) -> Tensor:
    """decoder.py:147-157."""
    x = F.interpolate(x, scale_factor=1.0 / factor, mode="linear")
    res = _conv(P, name + ".down_res", x)
    y = x
    for cname, d in (("c1", 1), ("c2", 2), ("c3", 4)):
        y = _conv(P, f"{name}.{cname}", F.leaky_relu(y, 0.1), dilation=d, replicate_pad=d)
    return y + res


def upsample_block(P: Params, name: str, x: Tensor, c: Tensor, factor: int) -> Tensor:
    """decoder.py:173-190."""
    x = F.interpolate(x, scale_factor=factor, mode="linear")
    for (ca, da), (cb, db), fl in ((("c1", 1), ("c2", 3), "film1"), (("c3", 9), ("c4", 27), "film2")):
        y = _conv(P, f"{name}.{ca}", F.leaky_relu(x, 0.1), dilation=da, replicate_pad=da)
        y = _conv(P, f"{name}.{cb}", F.leaky_relu(y, 0.1), dilation=db, replicate_pad=db)
        x = film(P, f"{name}.{fl}", y, c) + x
    return _conv(P, name + ".c5", x)


def filter_net(P: Params, content: Tensor, f0: Tensor, energy: Tensor, source: Tensor,
               prefix: str = "filter_net", return_skips: bool = False):
    """decoder.py:222-233."""
    x = _conv(P, prefix + ".content_in", content) + _conv(P, prefix + ".f0_in", _log_f0(f0))
    s = torch.cat([source, energy], dim=1)
    skips = []
    s = _conv(P, prefix + ".downs.0", s, replicate_pad=1)
    skips.append(s)
    for i, f in enumerate(reversed(FILTER_FACTORS[1:]), start=1):      # 5, 4, 4, 3
        s = downsample_block(P, f"{prefix}.downs.{i}", s, f)
        skips.append(s)
    for i, (f, c) in enumerate(zip(FILTER_FACTORS, reversed(skips))):  # x2 x3 x4 x4 x5
        x = upsample_block(P, f"{prefix}.ups.{i}", x, c, f)
    out = _conv(P, prefix + ".output_layer", x, replicate_pad=3)
    return (out, skips) if return_skips else out


def decoder_infer(P: Params, content: Tensor, f0: Tensor, energy: Tensor, rand01: Tensor,
                  return_parts: bool = False):
    """decoder.py:253-257 with the noise draw injected (see oscillate_noise)."""
    amps, kern = source_net(P, content, f0, energy)
    src = decoder_dsp(f0, amps, kern, rand01)
    out = filter_net(P, content, f0, energy, src).squeeze(1)
    if return_parts:
        return out, dict(amps=amps, kernel=kern, source=src)
    return out


# --------------------------------------------------------------------------------------------
# module/infer/generator.py
# --------------------------------------------------------------------------------------------
def generator_encode(PE: Params, wf: Tensor) -> Tuple[Tensor, Tensor]:
    """generator.py:18-23."""
    return encoder_infer(PE, spectrogram(autopad_waveform(wf)))


def generator_convert(PE: Params, PD: Params, wf: Tensor, tgt: Tensor, pitch_shift: float,
                      rand01: Optional[Tensor] = None, return_parts: bool = False):
    """generator.py:25-34.  `rand01` = the [B,961,Lf] uniform draw decoder.py:78 would make; if
    None it is drawn from torch's global CPU generator exactly as the reference does."""
    wf = autopad_waveform(wf)
    spec = spectrogram(wf)
    energy = estimate_energy(wf)
    z, f0 = encoder_infer(PE, spec)
    zm, idx = match_features(z, tgt, return_indices=True)
    f0s = shift_frequency(f0, pitch_shift)
    if rand01 is None:
        rand01 = torch.rand(wf.shape[0], FFT_BIN, spec.shape[2])
    out = decoder_infer(PD, zm, f0s, energy, rand01)
    if return_parts:
        return out, dict(spec=spec, energy=energy, z=z, f0=f0, idx=idx, zm=zm, f0s=f0s)
    return out


# --------------------------------------------------------------------------------------------
# module/infer/stream.py
# --------------------------------------------------------------------------------------------
def phase_vocoder(a: Tensor, b: Tensor, fade_out: Tensor, fade_in: Tensor) -> Tensor:
    """stream.py:9-26."""
    window = torch.sqrt(fade_out * fade_in)
    fa = torch.fft.rfft(a * window)
    fb = torch.fft.rfft(b * window)
    absab = torch.abs(fa) + torch.abs(fb)
    n = a.shape[0]
    if n % 2 == 0:
        absab[1:-1] *= 2
    else:
        absab[1:] *= 2
    phia = torch.angle(fa)
    dphi = torch.angle(fb) - phia
    dphi = dphi - 2 * math.pi * torch.floor(dphi / 2 / math.pi + 0.5)
    w = 2 * math.pi * torch.arange(n // 2 + 1).to(a) + dphi
    t = torch.arange(n).unsqueeze(-1).to(a) / n
    return a * (fade_out ** 2) + b * (fade_in ** 2) + torch.sum(absab * torch.cos(w * t + phia), -1) * window / n


class StreamOracle:
    """stream.py:30-95 -- rolling window + SOLA cross-fade around `generator_convert`.

    `convert_fn(window[1,T]) -> [1,T]` is injected so tests can drive the SOLA logic with the
    oracle's convert, the CUDA convert, or a stub."""

    def __init__(self, convert_fn, block_size: int = 1920, extra_size: int = 0,
                 use_phase_vocoder: bool = False):
        self.convert_fn = convert_fn
        self.block_size = block_size
        self.sola_search_size = 1920
        self.last_delay_size = 3840
        self.crossfade_size = 1920
        self.use_phase_vocoder = use_phase_vocoder
        self.input_size = max(block_size + self.crossfade_size + self.sola_search_size
                              + 2 * self.last_delay_size, block_size + extra_size)
        cf = self.crossfade_size
        self.fade_in = torch.sin(math.pi * torch.arange(0, 1, 1 / cf) / 2) ** 2
        self.fade_out = 1 - self.fade_in
        self.input_wav = torch.zeros(self.input_size)
        self.sola_buffer = torch.zeros(cf)
        self.last_shift = -1

    def audio_callback(self, block: Tensor) -> Tensor:
        bs, cf, ss, ld = self.block_size, self.crossfade_size, self.sola_search_size, self.last_delay_size
        self.input_wav = torch.roll(self.input_wav, -bs)
        self.input_wav[-bs:] = block
        y = self.convert_fn(self.input_wav[None]).squeeze(0).clone()
        temp = y[-bs - cf - ss - ld:-ld]
        ci = temp[None, None, :cf + ss]
        nom = F.conv1d(ci, self.sola_buffer[None, None, :])
        den = torch.sqrt(F.conv1d(ci ** 2, torch.ones(1, 1, cf)) + 1e-8)
        shift = int(torch.argmax(nom[0, 0] / den[0, 0]).item())
        self.last_shift = shift
        temp = temp[shift: shift + bs + cf]
        if self.use_phase_vocoder:
            temp[:cf] = phase_vocoder(self.sola_buffer, temp[:cf], self.fade_out, self.fade_in)
        else:
            temp[:cf] *= self.fade_in
            temp[:cf] += self.sola_buffer * self.fade_out
        self.sola_buffer = temp[-cf:]
        return temp[:-cf]
