"""Shared machinery: a torch module's parameters -> a native handle (packed once, re-packed on change)."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib


class NativeHandle:
    """Owns a tvc_decoder_t / tvc_encoder_t built from `module`'s current parameters."""

    def __init__(self, module: nn.Module, kind: int):
        self._module = module
        self._kind = kind
        self._h: Optional[ctypes.c_void_p] = None
        self._key: Optional[Tuple] = None
        self._device: Optional[torch.device] = None

    def _fingerprint(self) -> Tuple:
        def ver(p):
            try:
                return p._version
            except RuntimeError:    # inference tensors do not track a version counter
                return -1
        return tuple((p.data_ptr(), ver(p)) for p in self._module.parameters())

    def get(self) -> ctypes.c_void_p:
        key = self._fingerprint()
        if self._h is not None and key == self._key:
            return self._h
        self.release()
        L = _lib.lib()
        sd = self._module.state_dict()
        expect = _lib.param_names(self._kind)
        got = tuple((k, v.numel()) for k, v in sd.items())
        if got != expect:
            bad = next((i for i, (a, b) in enumerate(zip(got, expect)) if a != b), min(len(got), len(expect)))
            raise RuntimeError(
                f"state_dict does not match the native parameter table at entry {bad}: "
                f"module has {got[bad] if bad < len(got) else None}, library expects {expect[bad] if bad < len(expect) else None}")
        dev = next(self._module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError(f"{type(self._module).__name__} is on {dev}; tinyvc_b200 runs on CUDA only "
                               "(move it with .to('cuda')); there is no CPU path")
        flat = torch.cat([v.detach().reshape(-1).to(torch.float32) for v in sd.values()]).contiguous()
        out = ctypes.c_void_p()
        create = L.tvc_decoder_create if self._kind == _lib.KIND_DECODER else L.tvc_encoder_create
        with torch.cuda.device(dev):
            torch.cuda.current_stream(dev).synchronize()
            _lib.check(create(flat.data_ptr(), flat.numel(), ctypes.byref(out)), "weight upload")
        self._h, self._key, self._device = out, key, dev
        return out

    @property
    def device(self) -> torch.device:
        self.get()
        return self._device

    def release(self) -> None:
        if self._h is not None:
            L = _lib.lib()
            (L.tvc_decoder_destroy if self._kind == _lib.KIND_DECODER else L.tvc_encoder_destroy)(self._h)
            self._h, self._key = None, None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
