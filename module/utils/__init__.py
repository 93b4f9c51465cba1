from tinyvc_b200.utils import autopad_waveform, estimate_energy, resample, shift_frequency, spectrogram  # noqa: F401
