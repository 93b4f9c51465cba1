"""estimate_energy (reference utils/energy_estimation.py:9-14) via `tvc_estimate_energy`:
max-pool(|x|, 128, 64, pad 32) then linear interpolation back to the input length.  [B,L] -> [B,1,L]."""
from __future__ import annotations

import torch

from .. import _lib


@torch.inference_mode()
def estimate_energy(wave: torch.Tensor, frame_size: int = 64) -> torch.Tensor:
    if frame_size != 64:
        raise ValueError("the CUDA energy kernel is specialised for frame_size=64")
    wave = _lib.dev_f32(wave, "wave")
    if wave.dim() != 2:
        raise RuntimeError(f"estimate_energy: expected [B,L], got {tuple(wave.shape)}")
    B, L = wave.shape
    dev = wave.device
    Lib = _lib.lib()
    energy = torch.empty(B, 1, L, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        ws = _lib.WORKSPACE.get(Lib.tvc_energy_workspace_bytes(B, L), dev)
        _lib.check(Lib.tvc_estimate_energy(wave.data_ptr(), energy.data_ptr(), B, L, ws.data_ptr(), ws.numel(),
                                           _lib.stream_ptr(dev)), "tvc_estimate_energy")
    return energy
