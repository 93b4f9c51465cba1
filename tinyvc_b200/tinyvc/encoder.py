"""Drop-in for `module.tinyvc.encoder` (reference module/tinyvc/encoder.py) on sm_100a CUDA.

Same class names, constructor signatures, attributes and state_dict keys (so `encoder.pt` loads
strictly; infer.py:34-35, extract_index.py:27-28,45).  Forward passes go through
`tvc_encoder_forward` / `tvc_pitch_decode` (include/tinyvc_b200.h).
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ._native import NativeHandle
from .convnext import ConvNeXtLayer, LayerNorm

N_FFT = 1920
FFT_BIN = N_FFT // 2 + 1


def _owner_of(mod) -> "Encoder":
    owner = mod._owner() if getattr(mod, "_owner", None) is not None else None
    if owner is None:
        raise RuntimeError(f"{type(mod).__name__} must belong to a tinyvc_b200 Encoder: the native weight pack "
                           "covers both estimators")
    return owner


class PitchEstimator(nn.Module):
    """961 -> 128, LN, 4 ConvNeXt, -> 512 pitch-class logits; top-4 soft decode to Hz
    (reference encoder.py:11-72)."""

    def __init__(self, n_fft=1920, internal_channels=128, num_layers=4, num_classes=512, classes_per_octave=48,
                 min_frequency=20.0):
        super().__init__()
        if (n_fft, internal_channels, num_layers, num_classes, classes_per_octave, min_frequency) != \
                (N_FFT, 128, 4, 512, 48, 20.0):
            raise ValueError("the CUDA kernels are specialised for the reference's default PitchEstimator")
        self.num_classes, self.classes_per_octave, self.min_frequency = num_classes, classes_per_octave, min_frequency
        self.input_layer = nn.Conv1d(n_fft // 2 + 1, internal_channels, 1)
        self.norm = LayerNorm(internal_channels)
        self.mid_layers = nn.Sequential(*[ConvNeXtLayer(internal_channels) for _ in range(num_layers)])
        self.output_layer = nn.Conv1d(internal_channels, num_classes, 1)
        self._owner = None

    @torch.inference_mode()
    def forward(self, spec) -> torch.Tensor:
        return _owner_of(self)._run(spec, want_z=False, want_logits=True, want_f0=False)[1]

    # Host-side helpers kept for API parity (used by the reference's train_encoder.py:65,83); plain torch.
    def freq2id(self, f):
        return torch.ceil(torch.clamp(self.classes_per_octave * torch.log2(f / self.min_frequency), 0,
                                      self.num_classes - 1)).to(torch.long)

    def id2freq(self, ids):
        x = self.min_frequency * (2 ** (ids.to(torch.float) / self.classes_per_octave))
        x[x <= self.min_frequency] = 0
        return x

    @torch.inference_mode()
    def decode(self, logits, k: int = 4) -> torch.Tensor:
        """logits [B,512,Lf] -> f0 [B,1,Lf]  (reference encoder.py:61-67)."""
        if k != 4:
            raise ValueError("the CUDA pitch decoder is specialised for k=4")
        logits = _lib.dev_f32(logits, "logits")
        B, C, Lf = logits.shape
        if C != self.num_classes:
            raise RuntimeError(f"decode: expected [B,{self.num_classes},Lf], got {tuple(logits.shape)}")
        f0 = torch.empty(B, 1, Lf, device=logits.device, dtype=torch.float32)
        with torch.cuda.device(logits.device):
            _lib.check(_lib.lib().tvc_pitch_decode(logits.data_ptr(), f0.data_ptr(), B, Lf,
                                                   _lib.stream_ptr(logits.device)), "tvc_pitch_decode")
        return f0

    @torch.inference_mode()
    def infer(self, spec) -> torch.Tensor:
        return _owner_of(self)._run(spec, want_z=False, want_logits=False, want_f0=True)[2]


class SSLFeatureEstimator(nn.Module):
    """961 -> 384, LN, 6 ConvNeXt (dilations 1,3,9,1,1,1), -> 768 content features
    (reference encoder.py:75-97)."""

    def __init__(self, n_fft=1920, internal_channels=384, dilations=(1, 3, 9, 1, 1, 1), ssl_dim=768):
        super().__init__()
        if (n_fft, internal_channels, tuple(dilations), ssl_dim) != (N_FFT, 384, (1, 3, 9, 1, 1, 1), 768):
            raise ValueError("the CUDA kernels are specialised for the reference's default SSLFeatureEstimator")
        self.input_layer = nn.Conv1d(n_fft // 2 + 1, internal_channels, 1)
        self.norm = LayerNorm(internal_channels)
        self.mid_layers = nn.Sequential(*[ConvNeXtLayer(internal_channels, dilation=d) for d in dilations])
        self.output_layer = nn.Conv1d(internal_channels, ssl_dim, 1)
        self._owner = None

    @torch.inference_mode()
    def forward(self, spec) -> torch.Tensor:
        return _owner_of(self)._run(spec, want_z=True, want_logits=False, want_f0=False)[0]

    def infer(self, spec) -> torch.Tensor:
        return self.forward(spec)


class Encoder(nn.Module):
    """Reference `Encoder` (encoder.py:100-116)."""

    def __init__(self, n_fft=1920, hop_size=480):
        super().__init__()
        if (n_fft, hop_size) != (N_FFT, 480):
            raise ValueError("the CUDA kernels are specialised for n_fft=1920, hop_size=480")
        self.n_fft, self.hop_size = n_fft, hop_size
        self.ssl_feature_estimator = SSLFeatureEstimator(n_fft)
        self.pitch_estimator = PitchEstimator(n_fft)
        ref = weakref.ref(self)
        self.ssl_feature_estimator._owner = ref
        self.pitch_estimator._owner = ref
        self._native = NativeHandle(self, _lib.KIND_ENCODER)

    def _run(self, spec, want_z: bool, want_logits: bool, want_f0: bool):
        spec = _lib.dev_f32(spec, "spec")
        if spec.dim() != 3 or spec.shape[1] != FFT_BIN:
            raise RuntimeError(f"Encoder: expected spec [B,961,Lf], got {tuple(spec.shape)}")
        B, _, Lf = spec.shape
        dev = spec.device
        L = _lib.lib()
        h = self._native.get()
        z = torch.empty(B, 768, Lf, device=dev, dtype=torch.float32) if want_z else None
        logits = torch.empty(B, 512, Lf, device=dev, dtype=torch.float32) if want_logits else None
        f0 = torch.empty(B, 1, Lf, device=dev, dtype=torch.float32) if want_f0 else None
        with torch.cuda.device(dev):
            ws = _lib.WORKSPACE.get(L.tvc_encoder_workspace_bytes(B, Lf), dev)
            _lib.check(L.tvc_encoder_forward(h, spec.data_ptr(), _lib.ptr(z), _lib.ptr(logits), _lib.ptr(f0), B, Lf,
                                             ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)), "tvc_encoder_forward")
        return z, logits, f0

    @torch.inference_mode()
    def forward(self, spec) -> Tuple[torch.Tensor, torch.Tensor]:
        z, logits, _ = self._run(spec, True, True, False)
        return z, logits

    @torch.inference_mode()
    def infer(self, spec) -> Tuple[torch.Tensor, torch.Tensor]:
        z, _, f0 = self._run(spec, True, False, True)
        return z, f0
