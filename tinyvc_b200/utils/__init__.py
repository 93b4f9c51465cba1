"""Drop-in for the inference-time part of the reference's `module.utils`
(module/utils/__init__.py:1-5).

`estimate_f0` (pyworld / FCPE wrappers, f0_estimation.py) is training / preprocessing only and
is never called by `Generator.convert` (generator.py:26-34 ignores its `f0_estimation` argument),
so it is not provided; importing this package does not need pyworld or torchfcpe.
`resample` is the device-side counterpart of the `torchaudio.functional.resample` call in infer.py:45-46,63-64.
"""
from .auto_padding import autopad_waveform
from .energy_estimation import estimate_energy
from .pitch_shift import shift_frequency
from .resample import resample
from .spectrogram import spectrogram

__all__ = ["autopad_waveform", "estimate_energy", "resample", "shift_frequency", "spectrogram"]
