"""Parameter containers for the ConvNeXt-v2 block (drop-in for module/tinyvc/convnext.py).

The arithmetic of these layers runs inside the fused stack kernels driven from
`SourceNet`, `SSLFeatureEstimator` and `PitchEstimator` (csrc/nets.cu `convnext_forward`:
depth-wise conv + LayerNorm in one kernel, 1x1 + GELU, a deterministic GRN reduction, and a
1x1 conv with the GRN affine folded into its prologue and the residual into its epilogue).
The classes here exist so that `state_dict()` has exactly the reference's keys and shapes.
"""
from __future__ import annotations

import torch
import torch.nn as nn

_STANDALONE = ("{} holds parameters only; its arithmetic runs inside the owning network's CUDA "
               "stack kernels. Call the owning Encoder/Decoder (or its source_net / filter_net / "
               "ssl_feature_estimator / pitch_estimator) instead.")


class LayerNorm(nn.Module):
    """Channel LayerNorm parameters: gamma, beta [C]  (reference convnext.py:7-19)."""

    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        if eps != 1e-5:
            raise ValueError("the CUDA kernels are specialised for eps=1e-5")
        self.channels, self.eps = channels, eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))

    def forward(self, x):
        raise RuntimeError(_STANDALONE.format("LayerNorm"))


class GRN(nn.Module):
    """Global response norm parameters: beta, gamma [1,C,1]  (reference convnext.py:23-34)."""

    def __init__(self, channels: int, eps: float = 1e-6):
        super().__init__()
        if eps != 1e-6:
            raise ValueError("the CUDA kernels are specialised for eps=1e-6")
        self.eps = eps
        self.beta = nn.Parameter(torch.zeros(1, channels, 1))
        self.gamma = nn.Parameter(torch.zeros(1, channels, 1))

    def forward(self, x):
        raise RuntimeError(_STANDALONE.format("GRN"))


class ConvNeXtLayer(nn.Module):
    """c1 (depth-wise k, dilation d, replicate pad) / norm / c2 (C->mC) / grn / c3 (mC->C)
    (reference convnext.py:38-58)."""

    def __init__(self, channels: int = 512, kernel_size: int = 7, mlp_mul: int = 2, dilation: int = 1):
        super().__init__()
        if kernel_size != 7 or mlp_mul != 2:
            raise ValueError("the CUDA kernels are specialised for kernel_size=7, mlp_mul=2")
        pad = (kernel_size * dilation - dilation) // 2
        self.dilation = dilation
        self.c1 = nn.Conv1d(channels, channels, kernel_size, 1, pad, groups=channels, dilation=dilation,
                            padding_mode="replicate")
        self.norm = LayerNorm(channels)
        self.c2 = nn.Conv1d(channels, channels * mlp_mul, 1)
        self.grn = GRN(channels * mlp_mul)
        self.c3 = nn.Conv1d(channels * mlp_mul, channels, 1)

    def forward(self, x):
        raise RuntimeError(_STANDALONE.format("ConvNeXtLayer"))
