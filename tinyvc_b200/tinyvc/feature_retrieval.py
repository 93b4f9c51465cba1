"""Drop-in for `module.tinyvc.feature_retrieval.match_features` (reference feature_retrieval.py:15-33).

    match_features(source [B,768,Lf], reference [1|B,768,N], k=4, alpha=0.0, metrics='cos') -> [B,768,Lf]

The reference's `bmm` needs `reference` expanded to the batch and materialises [B,N,768] and
[B,Lf,N]; here the index is shared across the batch (a [1,768,N] tensor, which is what
`index.pt` holds, or an expanded view of one) and prepared once per tensor: a normalised copy in
kernel layout and a row-major copy for the gather are cached on the native side, keyed by the
tensor's storage.  Per-utterance indices ([B,768,N] with B distinct targets) are handled by
looping over utterances.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from .. import _lib

CONTENT = 768
_CACHE_MAX = 4


class _Index:
    def __init__(self, ref2d: torch.Tensor, metric: int):
        self.N = ref2d.shape[1]
        self.device = ref2d.device
        # The cache key contains the tensor's address.  Holding the tensor keeps that address from being handed to a
        # different index by the caching allocator while this entry is alive (a recycled pointer with the same shape
        # would otherwise hit the cache and silently match against the previous index).
        self.keep = ref2d
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            _lib.check(_lib.lib().tvc_index_create(ref2d.data_ptr(), self.N, metric, ctypes.byref(h)), "tvc_index_create")
        self.h = h

    def __del__(self):
        try:
            _lib.lib().tvc_index_destroy(self.h)
        except Exception:
            pass


_cache: "OrderedDict[Tuple, _Index]" = OrderedDict()


def _version(t: torch.Tensor) -> int:
    try:
        return t._version
    except RuntimeError:        # inference tensors do not track a version counter
        return -1


def _prepared(ref2d: torch.Tensor, metric: int) -> _Index:
    key = (ref2d.data_ptr(), _version(ref2d), ref2d.shape[1], metric, ref2d.device.index)
    hit = _cache.get(key)
    if hit is not None:
        _cache.move_to_end(key)
        return hit
    idx = _Index(ref2d, metric)
    _cache[key] = idx
    while len(_cache) > _CACHE_MAX:
        _cache.popitem(last=False)
    return idx


def clear_index_cache() -> None:
    _cache.clear()


def _match_shared(source, ref2d, k, alpha, metric, want_idx):
    B, _, Lf = source.shape
    dev = source.device
    index = _prepared(ref2d, metric)
    L = _lib.lib()
    out = torch.empty_like(source)
    idx = torch.empty(B, Lf, k, device=dev, dtype=torch.int32) if want_idx else None
    with torch.cuda.device(dev):
        ws = _lib.WORKSPACE.get(L.tvc_match_workspace_bytes(index.h, B, Lf), dev)
        _lib.check(L.tvc_match_features(index.h, source.data_ptr(), out.data_ptr(), _lib.ptr(idx), B, Lf, int(k),
                                        float(alpha), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                   "tvc_match_features")
    return out, idx


@torch.inference_mode()
def match_features(source, reference, k: int = 4, alpha: float = 0.0, metrics: str = "cos",
                   return_indices: bool = False):
    if metrics not in _lib.METRICS:
        raise ValueError(f"metrics must be one of {sorted(_lib.METRICS)}, got {metrics!r}")
    source = _lib.dev_f32(source, "source")
    if reference.dim() != 3 or source.dim() != 3 or source.shape[1] != CONTENT or reference.shape[1] != CONTENT:
        raise RuntimeError(f"match_features: expected source [B,768,Lf] and reference [B,768,N]; got "
                           f"{tuple(source.shape)}, {tuple(reference.shape)}")
    B = source.shape[0]
    if reference.shape[0] not in (1, B):
        raise RuntimeError(f"match_features: reference batch {reference.shape[0]} does not match source batch {B}")
    if not reference.is_cuda:
        raise RuntimeError("match_features: reference must be a CUDA tensor; there is no CPU path")
    metric = _lib.METRICS[metrics]
    shared = reference.shape[0] == 1 or reference.stride(0) == 0
    if shared:
        ref2d = _lib.dev_f32(reference[0], "reference")
        out, idx = _match_shared(source, ref2d, k, alpha, metric, return_indices)
    else:
        outs, idxs = [], []
        for b in range(B):
            o, i = _match_shared(source[b:b + 1].contiguous(), _lib.dev_f32(reference[b], "reference"), k, alpha, metric,
                                 return_indices)
            outs.append(o)
            idxs.append(i)
        out = torch.cat(outs, dim=0)
        idx = torch.cat(idxs, dim=0) if return_indices else None
    return (out, idx.to(torch.long)) if return_indices else out
