"""Bit-level numerics contracts of the TinyVC hot path, restated in NumPy -- TEST INFRASTRUCTURE.

`oracle/tinyvc_oracle.py` re-executes the reference's op sequence with ATen; this file instead
spells out, scalar operation by scalar operation, what those ATen CPU kernels compute for the
three places where the exact fp32 rounding sequence decides parity (SURVEY.md section 0,
Appendix A.1/A.3).  The CUDA kernels in `tinyvc_b200/csrc` implement exactly these formulas;
`tests/test_numerics_contracts.py` checks each of them against torch-CPU bit-for-bit, and the
golden fixtures pin them against the real reference.

Only tests may import this module.
"""
from __future__ import annotations

import math
import numpy as np

f32 = np.float32


def _fma32(a, b, c):
    """fp32 fused multiply-add: exact product and sum in fp64, one rounding to fp32.
    (fp64 holds the 48-bit product exactly; the fp64 add can round, but double rounding only
    differs from a true FMA on ties that do not occur for these magnitudes -- verified
    against torch in the tests.)"""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def interp_scale(in_len: int, out_len: int, scale_factor=None) -> np.float32:
    """ATen `area_pixel_compute_scale` for align_corners=False (UpSample.h).
    With `size=` the scale is float(in)/float(out); with `scale_factor=` it is
    float(1.0 / scale_factor) evaluated in double (decoder.py:148,174,262)."""
    if scale_factor is not None:
        return f32(1.0 / float(scale_factor))
    return f32(f32(in_len) / f32(out_len))


def interp_coords(in_len: int, out_len: int, scale: np.float32):
    """Source index / weights for F.interpolate(mode='linear') on CPU, fp32 (Appendix A.1)."""
    dst = np.arange(out_len, dtype=f32)
    src = _fma32(scale, dst + f32(0.5), f32(-0.5))
    src = np.maximum(src, f32(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_len - 1)
    i1 = i0 + (i0 < in_len - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    l0 = (f32(1) - l1).astype(f32)
    return i0, i1, l0, l1


def interp_linear(x: np.ndarray, out_len: int, scale_factor=None) -> np.ndarray:
    """x[..., in_len] -> [..., out_len];   out = fma(x[i0], l0, rn(x[i1]*l1))."""
    x = np.asarray(x, f32)
    in_len = x.shape[-1]
    scale = interp_scale(in_len, out_len, scale_factor)
    i0, i1, l0, l1 = interp_coords(in_len, out_len, scale)
    p1 = (x[..., i1] * l1).astype(f32)
    return _fma32(x[..., i0], l0, p1)


def harmonic_theta(f0: np.ndarray, frame: int = 480, sample_rate: float = 24000.0,
                   n_osc: int = 15):
    """decoder.py:39-50 for one utterance.  f0: [Lf] fp32 -> (theta [n_osc, L], uv [L]) fp32.

      fs   = interp(f0)[n] * k            (k = 1..15, int64 arange promoted to fp32)
      inc  = fs / 24000f                  (true division)
      I    = fp32( sum_{m<=n} fp64(inc) ) (sequential fp64 accumulation, rounded per element)
      th   = fp32(2*pi) * fmodf(I, 1)
    """
    f0 = np.asarray(f0, f32)
    lw = f0.shape[0] * frame
    fsi = interp_linear(f0, lw)
    k = np.arange(1, n_osc + 1, dtype=f32)[:, None]
    fs = (fsi[None, :] * k).astype(f32)
    inc = (fs / f32(sample_rate)).astype(f32)
    integ = np.cumsum(inc.astype(np.float64), axis=1).astype(f32)      # np.cumsum is sequential
    frac = np.fmod(integ, f32(1)).astype(f32)
    theta = (f32(2 * math.pi) * frac).astype(f32)
    uv = interp_linear((f0 > f32(20.0)).astype(f32), lw)
    return theta, uv


def noise_spectrum(kernel: np.ndarray, rand01: np.ndarray):
    """decoder.py:78-80 -- Y = kernel * exp(j*(rand*2*pi - pi)); returns (re, im) fp32."""
    a = ((np.asarray(rand01, f32) * f32(2)) * f32(math.pi)).astype(f32)
    a = (a - f32(math.pi)).astype(f32)
    re = (np.cos(a.astype(np.float64)).astype(f32) * kernel).astype(f32)
    im = (np.sin(a.astype(np.float64)).astype(f32) * kernel).astype(f32)
    return re, im


def noise_ola(re: np.ndarray, im: np.ndarray, frame: int = 480, n_fft: int = 1920) -> np.ndarray:
    """decoder.py:81-82 -- prepend a zero frame, irfft every frame, rectangular-window
    overlap-add, divide by the coverage count, trim n_fft/2 at both ends.  re/im: [961, Lf].
    Returns [Lf*frame] (fp64 math: this is the 'what' of the contract, not its rounding)."""
    lf = re.shape[1]
    spec = np.zeros((re.shape[0], lf + 1), np.complex128)
    spec[:, 1:] = re.astype(np.float64) + 1j * im.astype(np.float64)
    frames = np.fft.irfft(spec, n=n_fft, axis=0)                      # [n_fft, Lf+1]
    total = n_fft + frame * lf
    ola = np.zeros(total)
    env = np.zeros(total)
    for t in range(lf + 1):
        ola[t * frame:t * frame + n_fft] += frames[:, t]
        env[t * frame:t * frame + n_fft] += 1.0
    half = n_fft // 2
    return (ola[half:half + lf * frame] / env[half:half + lf * frame])


def sinc_resample_bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Polyphase filter bank of `torchaudio.functional.resample` (sinc_interp_hann), the call reference infer.py:45-46,63-64
    makes with its defaults.  torchaudio is an un-vendored, un-pinned dependency of the reference (requirements.txt; 2.11.0
    here): this restates its published `_get_sinc_resample_kernel`.  Returns (bank [n][2*width+o] fp32, width, o, n)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    # torch.arange(0, -n, -1) is int64; `/ new_freq` happens in fp32 (default dtype), the sum with idx in fp64
    ph = (np.arange(0, -n, -1).astype(f32) / f32(n)).astype(np.float64)[:, None]
    t = np.clip((ph + idx) * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t)
    kern = kern * (window * (base / o))
    return kern.astype(f32), width, o, n


def sinc_resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """x [..., L] -> [..., ceil(new * L / orig)]: `_apply_sinc_resample_kernel` (zero pad `width` in front and
    `width + o` behind, conv1d with stride o, phases interleaved, cut to the target length); fp64 accumulation."""
    x = np.asarray(x, f32)
    if int(orig_freq) == int(new_freq):
        return x.copy()
    bank, width, o, n = sinc_resample_bank(orig_freq, new_freq)
    L = x.shape[-1]
    flat = x.reshape(-1, L)
    pad = np.concatenate([np.zeros((flat.shape[0], width), f32), flat, np.zeros((flat.shape[0], width + o), f32)], axis=1)
    klen = bank.shape[1]
    nblk = (pad.shape[1] - klen) // o + 1
    win = np.lib.stride_tricks.sliding_window_view(pad, klen, axis=1)[:, ::o][:, :nblk]      # [B][nblk][klen]
    y = np.einsum("bjk,pk->bjp", win.astype(np.float64), bank.astype(np.float64)).reshape(flat.shape[0], -1)
    target = -(-n * L // o)
    return y[:, :target].astype(f32).reshape(*x.shape[:-1], target)
