"""spectrogram (reference utils/spectrogram.py:8-15) via `tvc_spectrogram`: centred reflect-padded
Hann-1920 / hop-480 STFT magnitude with frame 0 dropped, computed as a dense real-DFT product on
the GPU (no cuFFT).  wave [B,L] (L a multiple of 480) -> [B,961,L/480]."""
from __future__ import annotations

import torch

from .. import _lib


@torch.inference_mode()
def spectrogram(wave: torch.Tensor, n_fft: int = 1920, hop_size: int = 480) -> torch.Tensor:
    if (n_fft, hop_size) != (1920, 480):
        raise ValueError("the CUDA STFT is specialised for n_fft=1920, hop_size=480")
    dtype = wave.dtype
    wave = _lib.dev_f32(wave, "wave")
    if wave.dim() != 2:
        raise RuntimeError(f"spectrogram: expected [B,L], got {tuple(wave.shape)}")
    B, L = wave.shape
    if L % hop_size:
        raise RuntimeError(f"spectrogram: length {L} is not a multiple of {hop_size}; call autopad_waveform first")
    dev = wave.device
    Lib = _lib.lib()
    spec = torch.empty(B, n_fft // 2 + 1, L // hop_size, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        ws = _lib.WORKSPACE.get(Lib.tvc_spectrogram_workspace_bytes(B, L), dev)
        _lib.check(Lib.tvc_spectrogram(wave.data_ptr(), spec.data_ptr(), B, L, ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr(dev)), "tvc_spectrogram")
    return spec.to(dtype)
