"""autopad_waveform (reference utils/auto_padding.py:5-11): right zero-pad [B,T] to a multiple of 480."""
from __future__ import annotations

import torch


def autopad_waveform(wf: torch.Tensor, frame_size: int = 480) -> torch.Tensor:
    rem = wf.shape[1] % frame_size
    if rem == 0:
        return wf
    out = wf.new_zeros(wf.shape[0], wf.shape[1] + frame_size - rem)   # one allocation + one copy (device-side memcpy)
    out[:, : wf.shape[1]] = wf
    return out
