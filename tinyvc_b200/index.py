"""Building `index.pt`, the kNN database (reference extract_index.py:30-58).

The reference walks a shuffled `DataLoader(batch_size=1)` over the dataset cache, encodes one clip at a
time, keeps every `stride`-th content vector, stops once more than `size` vectors are collected, shuffles
the columns and saves the first `size` as a `[1, 768, size]` fp32 tensor.  Here the clips that the
reference would have visited are determined up front (same global-RNG draws, so the same clips and the
same column permutation under the same `torch.manual_seed`), then encoded in batches on the GPU.
"""
from __future__ import annotations

import os
from typing import Callable, List, Sequence

import torch

FRAME = 480


def dataloader_order(n: int) -> List[int]:
    """Visit order of `DataLoader(ds, batch_size=1, shuffle=True)` (extract_index.py:31) under the current global
    torch seed: the loader iterator draws its base seed, the RandomSampler draws its own seed, then permutes."""
    _base_seed = int(torch.empty((), dtype=torch.int64).random_().item())      # _BaseDataLoaderIter.__init__
    seed = int(torch.empty((), dtype=torch.int64).random_().item())            # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).tolist()


def columns_per_clip(num_samples: int, stride: int) -> int:
    """Content vectors a clip contributes: Lf = L // 480 frames (spectrogram.py:14 drops frame 0), every stride-th kept."""
    lf = num_samples // FRAME
    return (lf + stride - 1) // stride


def clips_needed(lengths: Sequence[int], order: Sequence[int], size: int, stride: int) -> List[int]:
    """Prefix of `order` the reference loop consumes: it stops after the clip that pushes the total past `size`
    (extract_index.py:47-51), or at the end of the dataset."""
    picked, total = [], 0
    for i in order:
        picked.append(i)
        total += columns_per_clip(lengths[i], stride)
        if total > size:
            break
    return picked


@torch.inference_mode()
def build_index(encoder, load_clip: Callable[[int], torch.Tensor], lengths: Sequence[int], size: int = 2048,
                stride: int = 4, batch_size: int = 64, device=None) -> torch.Tensor:
    """-> index [1, 768, min(size, available)] fp32 on the CPU (what `torch.save` writes, extract_index.py:58).

    `load_clip(i)` returns clip i as a mono waveform [L_i] (24 kHz); `lengths[i]` = L_i, a multiple of 480
    (the dataset cache holds fixed 48 000-sample clips, preprocess.py `-len`)."""
    from .utils import spectrogram
    device = torch.device(device) if device is not None else next(encoder.parameters()).device
    bad = [i for i, n in enumerate(lengths) if n % FRAME]
    if bad:
        raise RuntimeError(f"build_index: clip {bad[0]} has {lengths[bad[0]]} samples, not a multiple of {FRAME}")
    order = dataloader_order(len(lengths))
    picked = clips_needed(lengths, order, size, stride)
    feats: List[torch.Tensor] = [None] * len(picked)
    # batch clips of equal length; results go back to their visit position
    by_len = {}
    for pos, i in enumerate(picked):
        by_len.setdefault(lengths[i], []).append(pos)
    for n, positions in by_len.items():
        for o in range(0, len(positions), batch_size):
            chunk = positions[o:o + batch_size]
            wf = torch.stack([load_clip(picked[p]).reshape(-1).float() for p in chunk]).to(device)
            z, _ = encoder.infer(spectrogram(wf, encoder.n_fft, encoder.hop_size))
            z = z[:, :, ::stride].cpu()
            for j, p in enumerate(chunk):
                feats[p] = z[j:j + 1]
    features = torch.cat(feats, dim=2)
    perm = torch.randperm(features.size(2))                                    # extract_index.py:33-36,54
    return features.index_select(2, perm)[:, :, :size].contiguous()


def cache_lengths(dir_path: str) -> List[int]:
    """Sample counts of `<dir>/<idx>.wav` for idx = 0..n-1 (module/utils/dataset.py:6-20 layout)."""
    import torchaudio
    n = len([f for f in os.listdir(dir_path) if f.endswith(".wav")])
    out = []
    for i in range(n):
        info = torchaudio.info(os.path.join(dir_path, f"{i}.wav"))
        out.append(int(info.num_frames))
    return out


def cache_loader(dir_path: str) -> Callable[[int], torch.Tensor]:
    import torchaudio

    def load(i: int) -> torch.Tensor:
        wf, _ = torchaudio.load(os.path.join(dir_path, f"{i}.wav"))
        return wf.mean(dim=0)                                                  # dataset.py:17-18

    return load
