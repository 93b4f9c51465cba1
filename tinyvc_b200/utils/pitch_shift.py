"""shift_frequency (reference utils/pitch_shift.py:5-15) via `tvc_shift_frequency`: semitone shift
through log2(relu(f/440)+1e-6)*12+69, evaluated op-for-op in fp32 like the reference."""
from __future__ import annotations

import torch

from .. import _lib


@torch.inference_mode()
def shift_frequency(f0: torch.Tensor, shift: float) -> torch.Tensor:
    f0 = _lib.dev_f32(f0, "f0")
    out = torch.empty_like(f0)
    with torch.cuda.device(f0.device):
        _lib.check(_lib.lib().tvc_shift_frequency(f0.data_ptr(), out.data_ptr(), f0.numel(), float(shift),
                                                  _lib.stream_ptr(f0.device)), "tvc_shift_frequency")
    return out
