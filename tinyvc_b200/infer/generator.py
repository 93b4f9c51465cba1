"""Drop-in for `module.infer.generator.Generator` (reference module/infer/generator.py:12-34).

The conversion is organised as three device-side stages so that parity tests and the config-3 bench can time and
compare them one by one:

    analyse    waveform -> (padded waveform, spectrogram, energy, content z, f0)      front end + Encoder
    retarget   (z, f0)  -> (matched content, shifted f0)                                kNN over the index + pitch shift
    synthesise (content, f0, energy) -> waveform                                        Decoder

`convert` chains them and is the reference's public call.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from .. import tinyvc as _tv
from .. import utils as _ut


class Analysis(NamedTuple):
    wf: torch.Tensor        # [B, L]        input padded to whole 480-sample frames
    spec: torch.Tensor      # [B, 961, Lf]
    energy: torch.Tensor    # [B, 1, L]
    z: torch.Tensor         # [B, 768, Lf]
    f0: torch.Tensor        # [B, 1, Lf]


class Generator(nn.Module):
    def __init__(self, encoder: "_tv.Encoder", decoder: "_tv.Decoder"):
        super().__init__()
        self.encoder, self.decoder = encoder, decoder

    # ---- stages -----------------------------------------------------------------------------------
    def analyse(self, wf: torch.Tensor, with_energy: bool = True) -> Analysis:
        padded = _ut.autopad_waveform(wf)
        spec = _ut.spectrogram(padded)
        energy = _ut.estimate_energy(padded) if with_energy else None
        z, f0 = self.encoder.infer(spec)
        return Analysis(padded, spec, energy, z, f0)

    def retarget(self, z: torch.Tensor, f0: torch.Tensor, tgt: torch.Tensor, pitch_shift, want_indices: bool = False):
        if want_indices:
            zm, idx = _tv.match_features(z, tgt, return_indices=True)
        else:
            zm, idx = _tv.match_features(z, tgt), None
        return zm, _ut.shift_frequency(f0, pitch_shift), idx

    def synthesise(self, content, f0, energy, rand01: Optional[torch.Tensor] = None, keep=None) -> torch.Tensor:
        return self.decoder.infer(content, f0, energy, rand01=rand01, keep=keep)

    # ---- the reference's surface ------------------------------------------------------------------
    @torch.inference_mode()
    def encode(self, wf):
        """wf [B,T] -> (content [B,768,Lf], f0 [B,1,Lf])   (generator.py:18-23)."""
        a = self.analyse(wf, with_energy=False)
        return a.z, a.f0

    @torch.inference_mode()
    def convert(self, wf, tgt, pitch_shift, f0_estimation="default", device=torch.device("cpu"), *,
                rand01: Optional[torch.Tensor] = None, return_parts: bool = False, keep=None):
        """wf [B,T], tgt [1|B,768,N] -> waveform [B, ceil(T/480)*480]   (generator.py:25-34).

        `f0_estimation` and `device` are accepted and ignored, exactly like the reference (its
        convert never reads them; infer.py:66 even passes the device string in the f0 slot).
        Keyword-only extras: `rand01` injects the noise draw (see Decoder.infer);
        `return_parts` also returns the intermediates for stage-wise parity checks; `keep=(t0, t1)`: only those output
        samples will be read (Decoder.infer)."""
        a = self.analyse(wf)
        zm, f0s, idx = self.retarget(a.z, a.f0, tgt, pitch_shift, want_indices=return_parts)
        out = self.synthesise(zm, f0s, a.energy, rand01, keep)
        if return_parts:
            return out, dict(spec=a.spec, energy=a.energy, z=a.z, f0=a.f0, idx=idx, zm=zm, f0s=f0s)
        return out
