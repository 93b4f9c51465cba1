"""Drop-in for `module.tinyvc.decoder` (reference module/tinyvc/decoder.py) on hand-written sm_100a CUDA.

Same classes, constructor signatures, attribute names and state_dict keys as the reference, so
`decoder.pt` loads with `load_state_dict(strict=True)` and callers such as infer.py:36-37,66,
infer_streaming.py:41-43 and train_decoder.py:105-107 (`decoder.source_net`, `decoder.dsp`,
`decoder.filter_net`) keep working.  The arithmetic is not torch: every forward goes through the
C-ABI in include/tinyvc_b200.h (`tvc_decoder_infer`, `tvc_source_net`, `tvc_dsp`,
`tvc_filter_net`).  The nn.Conv1d objects below are parameter containers only.

Extensions over the reference signatures (keyword-only):
* `rand01=` carries the uniform [0,1) draw that the reference takes from torch's global generator inside
  `oscillate_noise` (decoder.py:78).  Passing the same tensor to the CPU reference (by seeding) and to this decoder makes
  the noise branch comparable.  When omitted, `infer` draws inside the noise kernel (Philox-4x32-10, a fresh tensor per
  call): the stream is seeded once per decoder from torch's CPU generator at first use, or explicitly with
  `Decoder.seed_noise(seed)` -- a later `torch.manual_seed` does NOT reseed it (call `seed_noise` for reproducible runs;
  `ShardedDecoder` gives every rank its own seed).  `dsp` / `filter_net` and the exact-fp32 plan draw `torch.rand` on the
  device instead.
* `out=` (infer only): destination buffer, possibly on a peer GPU or in pinned host memory.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ._native import NativeHandle
from .convnext import ConvNeXtLayer

FRAME = 480
N_FFT = 1920
FFT_BIN = N_FFT // 2 + 1
SAMPLE_RATE = 24000
NUM_HARMONICS = 14
CONTENT = 768

# a single native launch is kept under this much scratch; larger batches are split (utterances are independent)
MAX_WORKSPACE_BYTES = 16 << 30


def _need_owner(mod) -> "Decoder":
    owner = mod._owner() if getattr(mod, "_owner", None) is not None else None
    if owner is None:
        raise RuntimeError(f"{type(mod).__name__} must belong to a tinyvc_b200 Decoder: the native weight pack "
                           "covers the whole decoder (SourceNet + FilterNet)")
    return owner


class FiLM(nn.Module):
    """to_shift / to_scale 1x1 convs (reference decoder.py:88-97); fused into the conv epilogues."""

    def __init__(self, input_channels: int, condition_channels: int):
        super().__init__()
        self.to_shift = nn.Conv1d(condition_channels, input_channels, 1)
        self.to_scale = nn.Conv1d(condition_channels, input_channels, 1)


class Downsample(nn.Module):
    """Parameters of a FilterNet down block (reference decoder.py:137-157)."""

    def __init__(self, input_channels: int, output_channels: int, factor: int = 4):
        super().__init__()
        self.factor = factor
        self.down_res = nn.Conv1d(input_channels, output_channels, 1)
        for name, d, co in (("c1", 1, input_channels), ("c2", 2, input_channels), ("c3", 4, output_channels)):
            setattr(self, name, nn.Conv1d(input_channels, co, 3, 1, d, dilation=d, padding_mode="replicate"))


class Upsample(nn.Module):
    """Parameters of a FilterNet up block (reference decoder.py:160-190)."""

    def __init__(self, input_channels: int, output_channels: int, cond_channels: int, factor: int = 4):
        super().__init__()
        self.factor = factor
        c = input_channels

        def k3(d):
            return nn.Conv1d(c, c, 3, 1, d, dilation=d, padding_mode="replicate")

        self.c1, self.c2 = k3(1), k3(3)
        self.film1 = FiLM(c, cond_channels)
        self.c3, self.c4 = k3(9), k3(27)
        self.film2 = FiLM(c, cond_channels)
        self.c5 = nn.Conv1d(c, output_channels, 1)


class SourceNet(nn.Module):
    """content/f0/energy -> harmonic amplitudes [B,15,Lf] and noise filter [B,961,Lf]
    (reference decoder.py:102-134) via `tvc_source_net`."""

    def __init__(self, content_channels=768, channels=128, kernel_size=7, num_layers=3, n_fft=1920,
                 frame_size=480, num_harmonics=14, sample_rate=24000):
        super().__init__()
        if (content_channels, channels, kernel_size, num_layers, n_fft, frame_size, num_harmonics, sample_rate) != \
                (CONTENT, 128, 7, 3, N_FFT, FRAME, NUM_HARMONICS, SAMPLE_RATE):
            raise ValueError("the CUDA kernels are specialised for the reference's default SourceNet hyper-parameters")
        self.n_fft, self.frame_size = n_fft, frame_size
        self.num_harmonics, self.sample_rate = num_harmonics, sample_rate
        self.content_channels = content_channels
        self.content_in = nn.Conv1d(content_channels, channels, 1)
        self.energy_in = nn.Conv1d(1, channels, 1)
        self.f0_in = nn.Conv1d(1, channels, 1)
        self.mid_layers = nn.Sequential(*[ConvNeXtLayer(channels, kernel_size) for _ in range(num_layers)])
        self.to_amps = nn.Conv1d(channels, num_harmonics + 1, 1)
        self.to_kernel = nn.Conv1d(channels, n_fft // 2 + 1, 1)
        self._owner = None

    @torch.inference_mode()
    def forward(self, content, f0, energy) -> Tuple[torch.Tensor, torch.Tensor]:
        return _need_owner(self)._source_net(content, f0, energy)


class FilterNet(nn.Module):
    """U-Net of dilated convs turning the 16-channel source into the waveform
    (reference decoder.py:193-233) via `tvc_filter_net`."""

    def __init__(self, channels=(384, 192, 96, 48, 24), factors=(2, 3, 4, 4, 5), content_channels=768,
                 num_harmonics=14):
        super().__init__()
        channels, factors = list(channels), list(factors)
        if (channels, factors, content_channels, num_harmonics) != ([384, 192, 96, 48, 24], [2, 3, 4, 4, 5], CONTENT, 14):
            raise ValueError("the CUDA kernels are specialised for the reference's default FilterNet hyper-parameters")
        self.content_in = nn.Conv1d(content_channels, channels[0], 1)
        self.f0_in = nn.Conv1d(1, channels[0], 1)
        low_to_high = channels[::-1]                       # 24, 48, 96, 192, 384
        self.downs = nn.ModuleList([nn.Conv1d(num_harmonics + 3, low_to_high[0], 3, 1, 1, padding_mode="replicate")])
        for cin, cout, f in zip(low_to_high[:-1], low_to_high[1:], factors[:0:-1]):
            self.downs.append(Downsample(cin, cout, f))
        self.ups = nn.ModuleList(
            Upsample(cin, cout, cin, f) for cin, cout, f in zip(channels, channels[1:] + channels[-1:], factors))
        self.output_layer = nn.Conv1d(channels[-1], 1, 7, 1, 3, padding_mode="replicate")
        self._owner = None

    @torch.inference_mode()
    def forward(self, content, f0, energy, source) -> torch.Tensor:
        return _need_owner(self)._filter_net(content, f0, energy, source)


class Decoder(nn.Module):
    """Reference `Decoder` (decoder.py:236-266): infer(content, f0, energy) -> waveform [B, L]."""

    def __init__(self, sample_rate=24000, n_fft=1920, frame_size=480, num_harmonics=14):
        super().__init__()
        if (sample_rate, n_fft, frame_size, num_harmonics) != (SAMPLE_RATE, N_FFT, FRAME, NUM_HARMONICS):
            raise ValueError("the CUDA kernels are specialised for sample_rate=24000, n_fft=1920, frame_size=480, "
                             "num_harmonics=14")
        self.sample_rate, self.frame_size = sample_rate, frame_size
        self.num_harmonics, self.n_fft = num_harmonics, n_fft
        self.source_net = SourceNet(frame_size=frame_size, sample_rate=sample_rate, n_fft=n_fft)
        self.filter_net = FilterNet()
        ref = weakref.ref(self)
        self.source_net._owner = ref
        self.filter_net._owner = ref
        self._native = NativeHandle(self, _lib.KIND_DECODER)

    # ---- helpers ------------------------------------------------------------------------------
    def _prep(self, content, f0, energy):
        content = _lib.dev_f32(content, "content")
        f0 = _lib.dev_f32(f0, "f0")
        energy = _lib.dev_f32(energy, "energy")
        B, C, Lf = content.shape
        if C != CONTENT or tuple(f0.shape) != (B, 1, Lf) or tuple(energy.shape) != (B, 1, Lf * FRAME):
            raise RuntimeError(f"Decoder: expected content [B,768,Lf], f0 [B,1,Lf], energy [B,1,480*Lf]; got "
                               f"{tuple(content.shape)}, {tuple(f0.shape)}, {tuple(energy.shape)}")
        return content, f0, energy, B, Lf

    def seed_noise(self, seed: Optional[int] = None) -> None:
        """Seed the in-kernel noise draw used when `rand01` is not injected (the role torch.manual_seed plays for
        the reference's torch.rand, decoder.py:78).  Default: a value drawn from torch's CPU generator."""
        if seed is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
        h = self._native.get()
        dev = self._native.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tvc_decoder_seed(h, int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.stream_ptr(dev)), "tvc_decoder_seed")
        self._noise_seeded_for = h.value

    def _rand(self, rand01, B, Lf, device):
        if rand01 is None:
            return torch.rand(B, FFT_BIN, Lf, device=device)       # decoder.py:78
        rand01 = _lib.dev_f32(rand01, "rand01")
        if tuple(rand01.shape) != (B, FFT_BIN, Lf):
            raise RuntimeError(f"rand01: expected {(B, FFT_BIN, Lf)}, got {tuple(rand01.shape)}")
        return rand01

    @staticmethod
    def _batch_chunk(B: int, Lf: int) -> int:
        per1 = _lib.lib().tvc_decoder_infer_workspace_bytes(1, Lf)
        return max(1, min(B, int(MAX_WORKSPACE_BYTES // max(per1, 1))))

    # ---- reference API ------------------------------------------------------------------------
    @torch.inference_mode()
    def infer(self, content, f0, energy, *, rand01: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, keep: Optional[tuple] = None) -> torch.Tensor:
        """`keep=(t0, t1)` (extension): the caller will only read samples [t0, t1) of every waveform (a streaming tick keeps
        5 760 of its 13 440, module/infer/stream.py:75); the full-rate block then skips the windows that produce none of them.
        Samples inside the range are bit-identical to a full run; samples outside it are unspecified.
        `out` (extension): write the waveform [B, L] into this fp32 buffer instead of a fresh tensor.  It may live on
        a peer GPU that this device can address (a `tinyvc_b200.peer` window): the last kernel then stores straight
        into the peer's HBM over NVLink, which is how the sharded mode gathers without a copy.  It may also be a PINNED
        host tensor (`torch.empty(...).pin_memory()`): pinned memory is mapped into the device's address space, so the last
        kernel's 4-byte-per-sample stores travel over PCIe while it computes and no device-to-host copy follows
        (synchronise the stream before reading it on the host)."""
        content, f0, energy, B, Lf = self._prep(content, f0, energy)
        dev = content.device
        L = _lib.lib()
        h = self._native.get()
        if rand01 is not None or _lib.option("conv_impl", "tc") == "fp32":
            rand01 = self._rand(rand01, B, Lf, dev)      # the exact-fp32 plan has no in-kernel generator: torch.rand, like decoder.py:78
        elif getattr(self, "_noise_seeded_for", None) != h.value:
            self.seed_noise()            # the draw of decoder.py:78 happens inside the noise kernel
        if out is None:
            out = torch.empty(B, Lf * FRAME, device=dev, dtype=torch.float32)
        elif not ((out.is_cuda or out.is_pinned()) and out.dtype == torch.float32 and out.is_contiguous()
                  and tuple(out.shape) == (B, Lf * FRAME)):
            raise RuntimeError(f"Decoder.infer: out must be a contiguous fp32 tensor of shape {(B, Lf * FRAME)} on a CUDA "
                               "device or in pinned host memory")
        t0, t1 = (0, Lf * FRAME) if keep is None else (int(keep[0]), int(keep[1]))
        if not 0 <= t0 < t1 <= Lf * FRAME:
            raise RuntimeError(f"Decoder.infer: keep={keep} is not a non-empty range inside [0, {Lf * FRAME}]")
        step = self._batch_chunk(B, Lf)
        with torch.cuda.device(dev):
            for b0 in range(0, B, step):
                nb = min(step, B - b0)
                nbytes = L.tvc_decoder_infer_workspace_bytes(nb, Lf)
                ws = _lib.WORKSPACE.get(nbytes, dev)
                _lib.check(L.tvc_decoder_infer_range(h, content[b0:b0 + nb].data_ptr(), f0[b0:b0 + nb].data_ptr(),
                                                     energy[b0:b0 + nb].data_ptr(),
                                                     rand01[b0:b0 + nb].data_ptr() if rand01 is not None else None,
                                                     out[b0:b0 + nb].data_ptr(), nb, Lf, t0, t1, ws.data_ptr(), ws.numel(),
                                                     _lib.stream_ptr(dev)), "tvc_decoder_infer")
        return out

    @torch.inference_mode()
    def dsp(self, f0, amps, kernel, *, rand01: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B,16,L] = cat(harmonics * interp(amps), noise)  (reference decoder.py:259-266)."""
        f0, amps, kernel = _lib.dev_f32(f0, "f0"), _lib.dev_f32(amps, "amps"), _lib.dev_f32(kernel, "kernel")
        B, _, Lf = f0.shape
        if tuple(amps.shape) != (B, NUM_HARMONICS + 1, Lf) or tuple(kernel.shape) != (B, FFT_BIN, Lf):
            raise RuntimeError(f"dsp: expected amps [B,15,Lf], kernel [B,961,Lf]; got {tuple(amps.shape)}, {tuple(kernel.shape)}")
        dev = f0.device
        rand01 = self._rand(rand01, B, Lf, dev)
        L = _lib.lib()
        h = self._native.get()
        src = torch.empty(B, NUM_HARMONICS + 2, Lf * FRAME, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            ws = _lib.WORKSPACE.get(L.tvc_decoder_workspace_bytes(B, Lf), dev)
            _lib.check(L.tvc_dsp(h, f0.data_ptr(), amps.data_ptr(), kernel.data_ptr(), rand01.data_ptr(), src.data_ptr(),
                                 B, Lf, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)), "tvc_dsp")
        return src

    def forward(self, content, f0, energy, *, rand01: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.infer(content, f0, energy, rand01=rand01)

    # ---- sub-network entry points used by SourceNet / FilterNet ---------------------------------
    def _source_net(self, content, f0, energy):
        content, f0, energy, B, Lf = self._prep(content, f0, energy)
        dev = content.device
        L = _lib.lib()
        h = self._native.get()
        amps = torch.empty(B, NUM_HARMONICS + 1, Lf, device=dev, dtype=torch.float32)
        kern = torch.empty(B, FFT_BIN, Lf, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            ws = _lib.WORKSPACE.get(L.tvc_decoder_workspace_bytes(B, Lf), dev)
            _lib.check(L.tvc_source_net(h, content.data_ptr(), f0.data_ptr(), energy.data_ptr(), amps.data_ptr(),
                                        kern.data_ptr(), B, Lf, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                       "tvc_source_net")
        return amps, kern

    def _filter_net(self, content, f0, energy, source):
        content, f0, energy, B, Lf = self._prep(content, f0, energy)
        source = _lib.dev_f32(source, "source")
        if tuple(source.shape) != (B, NUM_HARMONICS + 2, Lf * FRAME):
            raise RuntimeError(f"filter_net: expected source [B,16,480*Lf], got {tuple(source.shape)}")
        dev = content.device
        L = _lib.lib()
        h = self._native.get()
        out = torch.empty(B, 1, Lf * FRAME, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            ws = _lib.WORKSPACE.get(L.tvc_decoder_workspace_bytes(B, Lf), dev)
            _lib.check(L.tvc_filter_net(h, content.data_ptr(), f0.data_ptr(), energy.data_ptr(), source.data_ptr(),
                                        out.data_ptr(), B, Lf, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                       "tvc_filter_net")
        return out


def harmonic_theta(f0: torch.Tensor) -> torch.Tensor:
    """Phase of `oscillate_harmonics` (reference decoder.py:39-50): f0 [B,1,Lf] -> theta [B,15,L].
    Parity probe for the fp64-scan contract."""
    f0 = _lib.dev_f32(f0, "f0")
    B, _, Lf = f0.shape
    theta = torch.empty(B, NUM_HARMONICS + 1, Lf * FRAME, device=f0.device, dtype=torch.float32)
    with torch.cuda.device(f0.device):
        _lib.check(_lib.lib().tvc_harmonic_theta(f0.data_ptr(), theta.data_ptr(), B, Lf, _lib.stream_ptr(f0.device)),
                   "tvc_harmonic_theta")
    return theta
