"""`torchaudio.functional.resample(waveform, orig_freq, new_freq)` as the reference calls it (infer.py:45-46,63-64: defaults
only -- sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) via `tvc_resample`: polyphase windowed-sinc filter bank built
once per (device, orig_freq, new_freq), one kernel per call.  The file is decoded on the host (torchaudio.load), the samples
go to the device as they are and everything after that -- resampling included -- runs there."""
from __future__ import annotations

import torch

from .. import _lib


@torch.inference_mode()
def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """waveform [..., L] (CUDA, fp32) -> [..., ceil(new_freq * L / orig_freq)]."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq:
        raise RuntimeError("resample: frequencies must be integers (torchaudio raises for non-integer rates too)")
    wf = _lib.dev_f32(waveform, "waveform")
    shape = wf.shape
    L = int(shape[-1])
    flat = wf.reshape(-1, L)
    B = flat.shape[0]
    lib = _lib.lib()
    Lout = int(lib.tvc_resample_length(L, int(orig_freq), int(new_freq)))
    if B == 0 or L == 0 or Lout < 0:
        raise RuntimeError(f"resample: invalid input shape {tuple(shape)} or rates {orig_freq} -> {new_freq}")
    out = torch.empty(B, Lout, device=wf.device, dtype=torch.float32)
    with torch.cuda.device(wf.device):
        _lib.check(lib.tvc_resample(flat.data_ptr(), out.data_ptr(), B, L, int(orig_freq), int(new_freq),
                                    _lib.stream_ptr(wf.device)), "tvc_resample")
    return out.reshape(*shape[:-1], Lout)
