"""Peer windows: a device tensor of one rank mapped into every rank's address space (CUDA IPC), one box.

The multi-GPU structure of this path is scatter -> convert -> gather with no reduction.  Over NVLink /
NVSwitch every GPU reaches every peer's HBM directly, so the ranks do not need a two-sided exchange at all: the owner
publishes its buffers once, the other ranks PULL their utterance blocks out of the owner's input tensors and PUSH
their waveforms into the owner's result tensor with plain device-to-device copies (`cudaMemcpyPeerAsync` behind
`Tensor.copy_`: copy engines, no SM is taken from the conv kernels, no NCCL kernel has to be co-scheduled with the
persistent tcgen05 kernels that fill every SM).  The process group is still needed, but only for control: shipping
the IPC handles once per buffer set and the two barriers around a step.

One process per GPU, all GPUs visible to every process (torchrun's default).  The handles are the ones
`torch.multiprocessing` uses to share CUDA tensors between processes (`reduce_tensor`), exchanged with
`broadcast_object_list`; opening a handle maps the owner's allocation into this process (on the owner's device
ordinal) and enables peer access between the two GPUs on first use.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor


def export_handles(tensors: Dict[str, torch.Tensor]) -> Dict[str, tuple]:
    """Picklable IPC descriptions of CUDA tensors (the owner must keep the tensors alive while peers use them)."""
    out = {}
    for k, t in tensors.items():
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError(f"peer window {k!r}: need a contiguous CUDA tensor")
        fn, args = reduce_tensor(t)
        out[k] = (fn, args)
    return out


def open_handles(handles: Dict[str, tuple]) -> Dict[str, torch.Tensor]:
    """Views of the owner's tensors in this process (device = the owner's ordinal)."""
    return {k: fn(*args) for k, (fn, args) in handles.items()}


def share_from(owner: int, tensors: Optional[Dict[str, torch.Tensor]], group=None) -> Dict[str, torch.Tensor]:
    """Collective: every rank gets a dict of tensors aliasing the owner's memory (the owner gets its own back)."""
    rank = dist.get_rank(group)
    box = [export_handles(tensors) if rank == owner else None]
    src = dist.get_global_rank(group, owner) if group is not None else owner
    dist.broadcast_object_list(box, src=src, group=group)
    if rank == owner:
        return dict(tensors)
    return open_handles(box[0])


class _RawCuda:
    """Minimal `__cuda_array_interface__` carrier: lets torch alias device memory it did not allocate."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class ResultWindow:
    """A [n, L] fp32 result buffer on the owner's GPU that every rank's KERNELS can store into.

    torch's IPC tensors are opened in a context of the owner's device, which is enough for copy engines but not for
    stores issued by another GPU's kernels; the library therefore allocates the buffer itself (`tvc_peer_alloc`) and
    every other rank maps it into ITS OWN device's address space (`tvc_peer_open`: `cudaIpcOpenMemHandle` with lazy peer
    access under the rank's device).  `tensor` aliases the memory as a torch tensor on this rank's device.
    Collective: construct on every rank of the group in the same order."""

    def __init__(self, owner: int, shape, device: torch.device, group=None):
        import ctypes

        from . import _lib
        self.device = torch.device(device)
        self.owner = dist.get_rank(group) == owner
        self._ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        n = 1
        for d in shape:
            n *= int(d)
        with torch.cuda.device(self.device):
            if self.owner:
                _lib.check(_lib.lib().tvc_peer_alloc(max(n, 1) * 4, ctypes.byref(self._ptr), handle), "tvc_peer_alloc")
            box = [bytes(handle) if self.owner else None]
            src = dist.get_global_rank(group, owner) if group is not None else owner
            dist.broadcast_object_list(box, src=src, group=group)
            if not self.owner:
                h = (ctypes.c_ubyte * 64).from_buffer_copy(box[0])
                _lib.check(_lib.lib().tvc_peer_open(h, ctypes.byref(self._ptr)), "tvc_peer_open")
            # torch labels the alias with the device it finds the memory on (the owner's ordinal); that label is only used
            # for slicing and copies -- kernels receive the raw address, which is valid under this rank's device
            self.tensor = torch.as_tensor(_RawCuda(self._ptr.value, shape))

    def close(self) -> None:
        from . import _lib
        if self._ptr:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                fn = _lib.lib().tvc_peer_free if self.owner else _lib.lib().tvc_peer_close
                fn(self._ptr)
            self._ptr = None
            self.tensor = None


def p2p_available(device: torch.device, owner_device: int) -> bool:
    """Whether kernels / copy engines of `device` can address the owner's GPU directly."""
    if device.type != "cuda":
        return False
    if device.index == owner_device:
        return True
    try:
        return bool(torch.cuda.can_device_access_peer(device.index, owner_device))
    except Exception:
        return False
