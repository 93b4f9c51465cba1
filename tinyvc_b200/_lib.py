"""ctypes binding of `libtinyvc_b200.so` (the C-ABI declared in include/tinyvc_b200.h).

This is the only place Python touches the native library.  There is no CPU fallback: if the
library is missing, or a tensor is not a CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p, POINTER
from typing import Dict, Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# TVC_LIB: developer override (same-box A/B of two builds of the library, tools/gpu_ab_lib.sh)
LIB_PATH = os.environ.get("TVC_LIB") or os.path.join(_HERE, "libtinyvc_b200.so")

_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()

KIND_DECODER = 0
KIND_ENCODER = 1
METRICS = {"cos": 0, "IP": 1, "L2": 2}

_P = c_void_p  # device pointers travel as plain integers

_SIGNATURES = {
    "tvc_last_error": (c_char_p, []),
    "tvc_version": (c_char_p, []),
    "tvc_set_option": (c_int, [c_char_p, c_char_p]),
    "tvc_launch_count": (ctypes.c_ulonglong, []),
    "tvc_measure_fp32_peak": (c_int, [ctypes.POINTER(ctypes.c_double), c_void_p]),
    "tvc_profile_report": (c_int, [c_char_p, c_size_t]),
    "tvc_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_ubyte)]),
    "tvc_peer_open": (c_int, [ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(c_void_p)]),
    "tvc_peer_close": (c_int, [c_void_p]),
    "tvc_peer_free": (c_int, [c_void_p]),
    "tvc_param_count": (c_int, [c_int]),
    "tvc_param_name": (c_char_p, [c_int, c_int]),
    "tvc_param_numel": (c_int64, [c_int, c_int]),
    "tvc_param_total": (c_int64, [c_int]),
    "tvc_decoder_create": (c_int, [_P, c_int64, POINTER(c_void_p)]),
    "tvc_decoder_destroy": (c_int, [c_void_p]),
    "tvc_decoder_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tvc_decoder_infer_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tvc_decoder_infer": (c_int, [c_void_p, _P, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_decoder_plan_windows": (c_int, [c_int, c_int64, c_int64, _P, _P]),
    "tvc_resample_length": (c_int64, [c_int64, c_int, c_int]),
    "tvc_resample": (c_int, [_P, _P, c_int, c_int64, c_int, c_int, c_void_p]),
    "tvc_decoder_infer_range": (c_int, [c_void_p, _P, _P, _P, _P, _P, c_int, c_int, c_int64, c_int64, _P, c_size_t, c_void_p]),
    "tvc_decoder_seed": (c_int, [c_void_p, ctypes.c_uint64, c_void_p]),
    "tvc_source_net": (c_int, [c_void_p, _P, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_dsp": (c_int, [c_void_p, _P, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_filter_net": (c_int, [c_void_p, _P, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_harmonic_theta": (c_int, [_P, _P, c_int, c_int, c_void_p]),
    "tvc_encoder_create": (c_int, [_P, c_int64, POINTER(c_void_p)]),
    "tvc_encoder_destroy": (c_int, [c_void_p]),
    "tvc_encoder_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tvc_encoder_forward": (c_int, [c_void_p, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_pitch_decode": (c_int, [_P, _P, c_int, c_int, c_void_p]),
    "tvc_index_create": (c_int, [_P, c_int, c_int, POINTER(c_void_p)]),
    "tvc_index_destroy": (c_int, [c_void_p]),
    "tvc_match_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "tvc_match_features": (c_int, [c_void_p, _P, _P, _P, c_int, c_int, c_int, c_float, _P, c_size_t, c_void_p]),
    "tvc_spectrogram_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tvc_spectrogram": (c_int, [_P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_energy_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tvc_estimate_energy": (c_int, [_P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_shift_frequency": (c_int, [_P, _P, c_int64, c_float, c_void_p]),
    "tvc_sola": (c_int, [_P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tvc_sola_pv_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tvc_sola_pv": (c_int, [_P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_phase_vocoder_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tvc_phase_vocoder": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, c_size_t, c_void_p]),
    "tvc_tc_conv_probe": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P,
                                  c_int, c_int, c_int, _P, _P, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load (once) and return the native library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"tinyvc_b200: native library not found at {LIB_PATH}. Build it with "
                        "`python -m tinyvc_b200.build` (needs nvcc); there is no CPU fallback.")
                h = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(h, name)   # AttributeError if the export is missing
                    fn.restype = res
                    fn.argtypes = args
                impl = os.environ.get("TVC_CONV_IMPL")
                if impl and h.tvc_set_option(b"conv_impl", impl.encode()) != 0:
                    raise RuntimeError(f"tinyvc_b200: TVC_CONV_IMPL={impl!r} is not a known conv implementation")
                if impl:
                    _options["conv_impl"] = impl
                for kv in filter(None, os.environ.get("TVC_OPTS", "").split(",")):      # e.g. TVC_OPTS=pdl=0,graphs=0
                    k, _, v = kv.partition("=")
                    if h.tvc_set_option(k.strip().encode(), v.strip().encode()) != 0:
                        raise RuntimeError(f"tinyvc_b200: TVC_OPTS entry {kv!r} was rejected by tvc_set_option")
                    _options[k.strip()] = v.strip()
                _lib = h
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().tvc_last_error()
        raise RuntimeError(f"{what} failed ({status}): {msg.decode() if msg else 'unknown error'}")


_options: dict = {}      # what this process has set through set_option / TVC_CONV_IMPL / TVC_OPTS (the library keeps the state)


def set_option(key: str, value: str) -> None:
    check(lib().tvc_set_option(key.encode(), value.encode()), f"tvc_set_option({key})")
    _options[key] = value


def option(key: str, default: str = "") -> str:
    """Last value this process gave `key` (set_option, TVC_CONV_IMPL, TVC_OPTS); `default` if it never set it."""
    lib()
    return _options.get(key, default)


def launch_count() -> int:
    return int(lib().tvc_launch_count())


def measure_fp32_peak(device=None) -> float:
    """Sustained CUDA-core FP32 FMA rate of the current device in TFLOP/s (FMA micro-benchmark inside the library)."""
    import torch
    v = ctypes.c_double(0.0)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    with torch.cuda.device(dev):
        check(lib().tvc_measure_fp32_peak(ctypes.byref(v), stream_ptr(dev)), "tvc_measure_fp32_peak")
    return float(v.value)


def profile_report() -> dict:
    """Per-launcher event times accumulated since the last call (needs set_option('profile','1'))."""
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().tvc_profile_report(buf, len(buf)), "tvc_profile_report")
    return json.loads(buf.value.decode())


def param_names(kind: int) -> Tuple[Tuple[str, int], ...]:
    L = lib()
    return tuple((L.tvc_param_name(kind, i).decode(), int(L.tvc_param_numel(kind, i)))
                 for i in range(L.tvc_param_count(kind)))


# ------------------------------------------------------------------------------------------------
# tensor plumbing
# ------------------------------------------------------------------------------------------------
def dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """Validate a kernel input: CUDA, fp32, contiguous (made so if needed)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tinyvc_b200 runs on CUDA only (got a {t.device} tensor); there is no CPU path")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Workspace:
    """Grow-only scratch buffer per (device, stream)."""

    def __init__(self) -> None:
        self._bufs: Dict[Tuple[int, int], torch.Tensor] = {}

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        key = (device.index if device.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream(device).cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            self._bufs.pop(key, None)
            buf = None
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def clear(self) -> None:
        self._bufs.clear()


WORKSPACE = Workspace()
