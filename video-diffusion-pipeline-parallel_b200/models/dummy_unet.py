"""DummyUNet: the stand-in model of the reference's CPU/gloo simulator (BASELINE config 1).

Same constructor, parameter names (``net.0``, ``net.2``, ``norm``) and arithmetic as reference
``src/models/dummy_unet.py:17-59``::

    out = x + tanh(step / 10) * Conv3d(SiLU(Conv3d(x))) + LayerNorm_C(x)

It is the baseline model, not the accelerated path: the simulator is CPU by definition.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class DummyUNet(nn.Module):
    def __init__(self, channels: int = 8, hidden_channels: int = 16,
                 use_layernorm: bool = True) -> None:
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv3d(channels, hidden_channels, kernel_size=3, padding=1),
            nn.SiLU(),
            nn.Conv3d(hidden_channels, channels, kernel_size=3, padding=1),
        )
        self.norm = nn.LayerNorm(channels) if use_layernorm else None

    def forward(self, latent: torch.Tensor, step: int) -> torch.Tensor:  # type: ignore[override]
        if latent.dim() < 2:
            raise ValueError("Latent tensor must have at least 2 dims (N, C, ...)")
        out = latent + math.tanh(step / 10.0) * self.net(latent)
        if self.norm is not None:
            # LayerNorm over the channel axis: move C last, normalise, move it back
            moved = latent.movedim(1, -1)
            normed = F.layer_norm(moved, self.norm.normalized_shape, self.norm.weight,
                                  self.norm.bias, self.norm.eps)
            out = out + normed.movedim(-1, 1)
        return out
