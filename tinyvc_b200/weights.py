"""Deterministic synthetic weights for TinyVC modules.

The pretrained `encoder.pt` / `decoder.pt` are not distributable with this repo (reference
README.md:10 points at Hugging Face) so parity tests and benchmarks run on random weights of the
reference's shapes.  The weights are a pure function of (key name, shape, seed): they do not
depend on module construction order, so the same call gives the same tensors for the
reference's classes (golden generation) and for this package's classes.

Scale follows torch's Conv1d default (uniform +-1/sqrt(fan_in) for weight and bias).  GRN
gamma/beta are drawn N(0, 0.5) instead of the reference's zeros (convnext.py:26-27), otherwise
every GRN would be the identity and its kernel would go untested (SURVEY.md 8c "Weights").
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) * 2654435761 + seed * 97 + 12345) % (2 ** 63))
    return g


def synth_state_dict(template: Mapping[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Return fp32 CPU tensors for every key of `template` (a state_dict)."""
    out: Dict[str, torch.Tensor] = {}
    fan_in: Dict[str, int] = {}
    for k, v in template.items():
        if k.endswith(".weight") and v.dim() == 3:
            fan_in[k[: -len(".weight")]] = v.shape[1] * v.shape[2]
    for k, v in template.items():
        g = _gen(k, seed)
        shape = tuple(v.shape)
        if ".grn." in k:
            t = torch.randn(shape, generator=g) * 0.5
        elif k.endswith("norm.gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("norm.beta"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            base = k.rsplit(".", 1)[0]
            bound = 1.0 / math.sqrt(fan_in.get(base, 1))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[k] = t.to(torch.float32)
    return out


def load_synth_weights(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """In-place: replace every parameter of `module` by its synthetic value."""
    sd = synth_state_dict(module.state_dict(), seed)
    dev = next(module.parameters()).device
    module.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    return module


def state_checksum(sd: Mapping[str, torch.Tensor]) -> float:
    """Order-independent fp64 checksum used to tie golden fixtures to a weight set."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].detach().to("cpu", torch.float64)
        tot += float(v.abs().sum()) + 0.5 * float(v.sum())
    return tot
