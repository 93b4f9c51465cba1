"""Drop-in for `module.infer.stream.StreamInfer` (reference module/infer/stream.py:30-95), plus a
batched variant for many concurrent streams on one GPU.

Per tick the reference rolls a 13 440-sample window, runs the whole `Generator.convert` on it,
finds the SOLA offset by normalised cross-correlation against the previous tail, cross-fades and
emits `block_size` samples (stream.py:68-95).  Here the window stays on the device, convert runs on
the CUDA path, and the SOLA search + cross-fade + tail update is one kernel (`tvc_sola`) with no
`.item()` host round trip.  Results depend on the 13 440-sample window exactly as in the reference
(GRN normalises over the window).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .. import _lib
from .generator import Generator


@torch.inference_mode()
def phase_vocoder(a: torch.Tensor, b: torch.Tensor, fade_out: torch.Tensor, fade_in: torch.Tensor) -> torch.Tensor:
    """Reference `phase_vocoder(a, b, fade_out, fade_in)` (stream.py:9-26) on the GPU; a, b [n] or [S,n].
    `fade_out` must be `1 - fade_in` (it always is: stream.py:61-62); only `fade_in` is passed down.
    For an all-zero `a` (the first tick after init_buffer) the reference's result depends on the signed zeros its
    FFT library returns (angle(-0.0) = pi); here angle(0) = 0 for every bin."""
    a2, b2 = _lib.dev_f32(a, "a").reshape(-1, a.shape[-1]), _lib.dev_f32(b, "b").reshape(-1, b.shape[-1])
    fade_in = _lib.dev_f32(fade_in, "fade_in")
    if a2.shape != b2.shape or fade_in.numel() != a2.shape[1] or fade_out.numel() != a2.shape[1]:
        raise RuntimeError(f"phase_vocoder: shapes {tuple(a.shape)}, {tuple(b.shape)}, {tuple(fade_out.shape)}, {tuple(fade_in.shape)}")
    S, n = a2.shape
    out = torch.empty_like(a2)
    L = _lib.lib()
    with torch.cuda.device(a2.device):
        ws = _lib.WORKSPACE.get(L.tvc_phase_vocoder_workspace_bytes(S, n), a2.device)
        _lib.check(L.tvc_phase_vocoder(a2.data_ptr(), b2.data_ptr(), fade_in.data_ptr(), out.data_ptr(), S, n, ws.data_ptr(),
                                       ws.numel(), _lib.stream_ptr(a2.device)), "tvc_phase_vocoder")
    return out.reshape(a.shape)


class BatchedStreamInfer:
    """S independent streams sharing one target index; state tensors are [S, ...] on `device`."""

    def __init__(self, generator: Generator, num_streams: int, target=None, pitch_shift: float = 0.0,
                 device=torch.device("cuda"), block_size: int = 1920, extra_size: int = 0,
                 use_phase_vocoder: bool = False, f0_estimation: str = "default"):
        self.generator = generator
        self.num_streams = int(num_streams)
        self.target = target
        self.pitch_shift = pitch_shift
        self.device = torch.device(device)
        self.block_size = block_size
        self.extra_size = extra_size
        self.sola_search_size = 1920
        self.last_dilay_size = 3840          # (sic) attribute name kept from the reference, stream.py:48
        self.crossfade_size = 1920
        self.use_phase_vocoder = use_phase_vocoder
        self.f0_estimation = f0_estimation
        self.input_size = max(self.block_size + self.crossfade_size + self.sola_search_size + 2 * self.last_dilay_size,
                              self.block_size + self.extra_size)
        self.last_shift: Optional[torch.Tensor] = None
        self.prune_output = True             # compute only the waveform samples the SOLA step reads (bit-identical there)
        self.use_graph = True                # replay the whole tick as one CUDA graph from the second tick on
        self._graph = None
        self._graph_key = None
        self._ticks = 0

    def init_buffer(self) -> None:
        if self.device.type != "cuda":
            raise RuntimeError(f"StreamInfer: device {self.device} is not CUDA; there is no CPU path")
        cf = self.crossfade_size
        # stream.py:61-62 (built once on the host side of the device; not on the per-tick path)
        self.fade_in_window = (torch.sin(math.pi * torch.arange(0, 1, 1 / cf, device=self.device) / 2) ** 2).contiguous()
        self.fade_out_window = 1 - self.fade_in_window
        self.input_wav = torch.zeros(self.num_streams, self.input_size, device=self.device)
        self.sola_buffer = torch.zeros(self.num_streams, cf, device=self.device)
        self._shift = torch.zeros(self.num_streams, dtype=torch.int32, device=self.device)
        self._graph, self._graph_key, self._ticks = None, None, 0

    # ---- one tick ---------------------------------------------------------------------------------------------------
    def _keep_range(self):
        # stream.py:75 reads y[-(block+cross+search+delay) : -delay] and nothing else: tell the decoder (exact pruning)
        Ly = -(-self.input_size // 480) * 480
        keep = (max(0, Ly - self.block_size - self.crossfade_size - self.sola_search_size - self.last_dilay_size),
                Ly - self.last_dilay_size)
        return keep if self.prune_output and keep[1] > keep[0] else None

    def _convert_and_sola(self, window: torch.Tensor, out: torch.Tensor, rand01) -> None:
        S, bs = self.num_streams, self.block_size
        y = self.generator.convert(window, self.target, self.pitch_shift, rand01=rand01, keep=self._keep_range()).contiguous()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            if self.use_phase_vocoder:        # stream.py:83-89
                ws = _lib.WORKSPACE.get(L.tvc_sola_pv_workspace_bytes(S, self.crossfade_size), self.device)
                _lib.check(L.tvc_sola_pv(y.data_ptr(), y.shape[1], self.sola_buffer.data_ptr(), self.fade_in_window.data_ptr(),
                                         out.data_ptr(), self._shift.data_ptr(), S, bs, self.crossfade_size,
                                         self.sola_search_size, self.last_dilay_size, ws.data_ptr(), ws.numel(),
                                         _lib.stream_ptr(self.device)), "tvc_sola_pv")
            else:                             # stream.py:90-92
                _lib.check(L.tvc_sola(y.data_ptr(), y.shape[1], self.sola_buffer.data_ptr(),
                                      self.fade_in_window.data_ptr(), out.data_ptr(), self._shift.data_ptr(), S, bs,
                                      self.crossfade_size, self.sola_search_size, self.last_dilay_size,
                                      _lib.stream_ptr(self.device)), "tvc_sola")

    def _graph_state(self):
        tgt = self.target
        return (tgt.data_ptr(), getattr(tgt, "_version", 0) if not tgt.is_inference() else -1, tuple(tgt.shape),
                float(self.pitch_shift), bool(self.use_phase_vocoder), bool(self.prune_output))

    def _tick_graph(self, blocks: torch.Tensor) -> torch.Tensor:
        """The whole tick -- window slide, analyse, retarget, synthesise, SOLA -- as ONE CUDA graph per (S, window, target):
        a tick then costs one graph launch on the host (a single stream's tick is host-bound otherwise: ~120 launches).
        State (window, SOLA tail, noise generator) lives in fixed device buffers that the graph updates in place."""
        n, bs = self.input_size, self.block_size
        state = self._graph_state()
        if self._graph is None or self._graph_key != state:
            from ..tinyvc.feature_retrieval import _prepared
            self._graph_index = _prepared(_lib.dev_f32(self.target[0], "reference"), _lib.METRICS["cos"])   # keep the native index alive
            self._g_blocks = torch.empty_like(blocks)
            self._g_slide = torch.empty(self.num_streams, n - bs, device=self.device)
            self._g_out = torch.empty(self.num_streams, bs, device=self.device)
            self._g_blocks.copy_(blocks)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                # stream.py:69-70, in place (two copies: the ranges overlap)
                self._g_slide.copy_(self.input_wav[:, bs:])
                self.input_wav[:, : n - bs].copy_(self._g_slide)
                self.input_wav[:, n - bs:].copy_(self._g_blocks)
                self._convert_and_sola(self.input_wav, self._g_out, None)
            self._graph, self._graph_key = g, state
        else:
            self._g_blocks.copy_(blocks)
        self._graph.replay()
        self.last_shift = self._shift
        return self._g_out.clone()

    @torch.inference_mode()
    def audio_callback(self, blocks: torch.Tensor, *, rand01: Optional[torch.Tensor] = None) -> torch.Tensor:
        """blocks [S, block_size] -> converted [S, block_size]."""
        S, bs = self.num_streams, self.block_size
        blocks = _lib.dev_f32(blocks, "block")
        if tuple(blocks.shape) != (S, bs):
            raise RuntimeError(f"audio_callback: expected {(S, bs)}, got {tuple(blocks.shape)}")
        # graph mode: from the second tick on (the first one runs eagerly: it triggers every one-time set-up -- kernel
        # attributes, the native index, workspaces -- none of which may happen inside a capture); injected noise draws
        # (parity tests) and per-utterance targets stay on the eager path
        if (self.use_graph and rand01 is None and self._ticks > 0 and isinstance(self.target, torch.Tensor)
                and self.target.dim() == 3 and self.target.shape[0] == 1):
            self._ticks += 1
            return self._tick_graph(blocks)
        self._ticks += 1
        # stream.py:69-70: slide the window left by one block and append the new samples -- in place, like the graph path, so
        # that eager and replayed ticks of one object can alternate (an injected rand01 takes this path) on the same state
        keep_part = self.input_wav[:, bs:].clone()
        self.input_wav[:, : self.input_size - bs] = keep_part
        self.input_wav[:, self.input_size - bs:] = blocks
        out = torch.empty(S, bs, device=self.device, dtype=torch.float32)
        self._convert_and_sola(self.input_wav, out, rand01)
        self.last_shift = self._shift
        return out


class StreamInfer(BatchedStreamInfer):
    """Single-stream API of the reference: audio_callback(block [block_size]) -> [block_size]."""

    def __init__(self, generator: Generator, target=None, pitch_shift: float = 0.0, device=torch.device("cpu"),
                 block_size: int = 1920, extra_size: int = 0, use_phase_vocoder: bool = False,
                 f0_estimation: str = "default"):
        super().__init__(generator, 1, target, pitch_shift, device, block_size, extra_size, use_phase_vocoder,
                         f0_estimation)

    @torch.inference_mode()
    def audio_callback(self, block: torch.Tensor, *, rand01: Optional[torch.Tensor] = None) -> torch.Tensor:
        return super().audio_callback(block.reshape(1, -1), rand01=rand01)[0]
