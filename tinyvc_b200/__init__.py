"""tinyvc_b200 -- B200-native (sm_100a) implementation of the TinyVC real-time inference path.

Sub-packages mirror the reference's `module` package so existing call sites keep working:

    tinyvc_b200.tinyvc   <->  module.tinyvc   (Encoder, Decoder, match_features)
    tinyvc_b200.infer    <->  module.infer    (Generator, StreamInfer)
    tinyvc_b200.utils    <->  module.utils    (autopad_waveform, spectrogram, estimate_energy, shift_frequency)

The repo-root `module/` package re-exports them under the reference's import paths.
All arithmetic runs in `libtinyvc_b200.so` (csrc/, C-ABI in include/tinyvc_b200.h); importing
this package does not load the library, calling any op does, and fails loudly if it is missing.
"""
__version__ = "0.1.0"
