"""Drop-in for `module.infer.generator.Generator` (reference module/infer/generator.py:12-34)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ..tinyvc import Decoder, Encoder, match_features
from ..utils import autopad_waveform, estimate_energy, shift_frequency, spectrogram


class Generator(nn.Module):
    def __init__(self, encoder: Encoder, decoder: Decoder):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder

    @torch.inference_mode()
    def encode(self, wf):
        """wf [B,T] -> (content [B,768,Lf], f0 [B,1,Lf])   (generator.py:18-23)."""
        spec = spectrogram(autopad_waveform(wf))
        return self.encoder.infer(spec)

    @torch.inference_mode()
    def convert(self, wf, tgt, pitch_shift, f0_estimation="default", device=torch.device("cpu"), *,
                rand01: Optional[torch.Tensor] = None, return_parts: bool = False):
        """wf [B,T], tgt [1|B,768,N] -> waveform [B, ceil(T/480)*480]   (generator.py:25-34).

        `f0_estimation` and `device` are accepted and ignored, exactly like the reference (its
        convert never reads them; infer.py:66 even passes the device string in the f0 slot).
        Keyword-only extras: `rand01` injects the noise draw (see Decoder.infer);
        `return_parts` also returns the intermediates for stage-wise parity checks."""
        wf = autopad_waveform(wf)
        spec = spectrogram(wf)
        energy = estimate_energy(wf)
        z, f0 = self.encoder.infer(spec)
        if return_parts:
            zm, idx = match_features(z, tgt, return_indices=True)
        else:
            zm, idx = match_features(z, tgt), None
        f0s = shift_frequency(f0, pitch_shift)
        out = self.decoder.infer(zm, f0s, energy, rand01=rand01)
        if return_parts:
            return out, dict(spec=spec, energy=energy, z=z, f0=f0, idx=idx, zm=zm, f0s=f0s)
        return out
