"""DummyUNet: the stand-in model of the reference's CPU/gloo simulator (BASELINE config 1).

Same constructor, parameter names (``net.0``, ``net.2``, ``norm``) and arithmetic as reference
``src/models/dummy_unet.py:17-59``::

    out = x + tanh(step / 10) * Conv3d(SiLU(Conv3d(x))) + LayerNorm_C(x)

It is the baseline model, not the accelerated path: the simulator is CPU by definition.  For an fp32 CUDA latent
(the reference's ``--model dummy`` GPU benchmark, benchmark.py:77-83) the step runs on ``svdpp_dummy_unet_step``
(direct 3x3x3 convolutions + fused combine, three launches, parity-tested against this module's torch arithmetic);
other dtypes and devices use the torch ops below.  ``native_cuda = False`` forces torch everywhere.
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
        self.native_cuda = True

    def _forward_native(self, latent: torch.Tensor, step: int) -> torch.Tensor:
        from .. import native
        w1, w2 = self.net[0].weight, self.net[2].weight
        x = latent.contiguous()
        hidden = torch.empty((x.shape[0], w1.shape[0]) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
        # the kernel adds LayerNorm_C(x) iff gamma is given
        g, b, eps = (self.norm.weight, self.norm.bias, self.norm.eps) if self.norm is not None else (None, None, 0.0)
        return native.dummy_unet_step(torch.empty_like(x), x, w1.contiguous(), self.net[0].bias, w2.contiguous(),
                                      self.net[2].bias, g, b, eps, math.tanh(step / 10.0), hidden)

    def forward(self, latent: torch.Tensor, step: int) -> torch.Tensor:  # type: ignore[override]
        if latent.dim() < 2:
            raise ValueError("Latent tensor must have at least 2 dims (N, C, ...)")
        if (self.native_cuda and latent.is_cuda and latent.dtype == torch.float32 and latent.dim() == 5
                and self.net[0].weight.dtype == torch.float32 and not torch.is_grad_enabled()):
            return self._forward_native(latent, step)
        out = latent + math.tanh(step / 10.0) * self.net(latent)
        if self.norm is not None:
            # LayerNorm over the channel axis: move C last, normalise, move it back
            moved = latent.movedim(1, -1)
            normed = F.layer_norm(moved, self.norm.normalized_shape, self.norm.weight,
                                  self.norm.bias, self.norm.eps)
            out = out + normed.movedim(-1, 1)
        return out
