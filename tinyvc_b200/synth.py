"""Seeded synthetic inputs of the shapes BASELINE.json names.

All tensors are generated on the CPU with a private torch.Generator so GPU runs, the CPU oracle
and the golden fixtures see the same values.  `rand01` is drawn exactly as `torch.rand` would
draw it, so it doubles as the reference's internal noise draw (decoder.py:78).
"""
from __future__ import annotations

from typing import Dict

import torch

FRAME = 480
FFT_BIN = 961
CONTENT = 768


def synth_f0(batch: int, lf: int, g: torch.Generator) -> torch.Tensor:
    """Random-walk pitch contour, 80..800 Hz, ~25 % unvoiced in runs of 5..30 frames. [B,1,Lf]"""
    walk = torch.cumsum(0.03 * torch.randn(batch, lf, generator=g), dim=1)
    f0 = (220.0 * torch.pow(torch.tensor(2.0), walk)).clamp(80.0, 800.0)
    for b in range(batch):
        t = int(torch.randint(0, 40, (1,), generator=g))
        while t < lf:
            run = int(torch.randint(5, 31, (1,), generator=g))
            f0[b, t:t + run] = 0.0
            t += run + int(torch.randint(30, 150, (1,), generator=g))
    return f0.unsqueeze(1).contiguous()


def decoder_inputs(batch: int, lf: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """content [B,768,Lf] ~N(0,1); f0 [B,1,Lf]; energy [B,1,L] ~U(0,1); rand01 [B,961,Lf] ~U(0,1)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    content = torch.randn(batch, CONTENT, lf, generator=g)
    f0 = synth_f0(batch, lf, g)
    energy = torch.rand(batch, 1, lf * FRAME, generator=g)
    rand01 = torch.rand(batch, FFT_BIN, lf, generator=g)
    return dict(content=content, f0=f0, energy=energy, rand01=rand01)


def pipeline_inputs(batch: int, samples: int, index_size: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """wf [B,T] = 0.1*N(0,1); index [1,768,N] ~N(0,1) (distinct columns)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    wf = 0.1 * torch.randn(batch, samples, generator=g)
    index = torch.randn(1, CONTENT, index_size, generator=g)
    return dict(wf=wf, index=index)
