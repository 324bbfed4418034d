"""ORACLE (test infrastructure, not product code): plain-torch restatement of diffusers
``AutoencoderKLTemporalDecoder`` - the VAE of Stable Video Diffusion that the reference uses to turn the conditioning
image into ``image_latents`` (``scripts/generate_video_demo.py:119-143``: ``vae.encode(x).latent_dist.mode()``) and the
final latents into frames (``:154-195``: ``vae.decode(chunk, num_frames=n).sample`` on chunks of frames, fp32 because
``force_upcast`` is set).  SURVEY.md section 8(f) rank 3.

diffusers is not installed here and its source is not under /root/reference, so this file restates the published
architecture from memory.  PARITY UNPINNED: the reference holds no vector for it; module / parameter names follow the
diffusers ``state_dict`` keys (``encoder.down_blocks.0.resnets.0.conv1.weight``, ``decoder.up_blocks.1.resnets.2.
temporal_res_block.conv1.weight``, ``decoder.time_conv_out.weight`` ...) so a real checkpoint can be loaded later, and
``tools/verify_against_diffusers.py`` compares against the real class when it becomes importable.

Choices that cannot be checked against the source here (UNVERIFIED):
  V1  GroupNorm eps 1e-6 everywhere in the encoder; decoder SpatioTemporalResBlocks: spatial 1e-6, temporal 1e-5
  V2  decoder AlphaBlender: merge_strategy "learned" (a scalar), switch_spatial_to_temporal_mix = True:
      out = (1 - sigmoid(mix)) * spatial + sigmoid(mix) * temporal, mix initialised to 0
  V3  Downsample2D of the encoder pads (0, 1, 0, 1) and convolves with stride 2, padding 0
  V4  mid-block attention: one head of width 512, GroupNorm(32, eps 1e-6) on the input, biased q/k/v/out projections,
      residual connection, no rescale
  V5  the temporal decoder has no post_quant_conv; the encoder's moments go through quant_conv (1x1, 8 -> 8)
  V6  decoder output: conv_norm_out (eps 1e-6) -> SiLU -> conv_out -> Conv3d (3,1,1) over frames (time_conv_out)

Only tests/, tools/ and __graft_entry__.smoke() may import this.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None."""

    def __init__(self, cin: int, cout: int, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class TemporalResnetBlock(nn.Module):
    def __init__(self, c: int, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, c, eps=eps)
        self.conv1 = nn.Conv3d(c, c, (3, 1, 1), padding=(1, 0, 0))
        self.norm2 = nn.GroupNorm(32, c, eps=eps)
        self.conv2 = nn.Conv3d(c, c, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x):  # [B, C, F, H, W]
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return x + h


class AlphaBlender(nn.Module):
    def __init__(self, alpha: float, switch: bool):
        super().__init__()
        self.mix_factor = nn.Parameter(torch.tensor([alpha]))
        self.switch = switch

    def forward(self, x_spatial, x_temporal):
        a = torch.sigmoid(self.mix_factor).to(x_spatial.dtype)
        if self.switch:
            a = 1.0 - a
        return a * x_spatial + (1.0 - a) * x_temporal


class SpatioTemporalResBlock(nn.Module):
    """Decoder flavour: no time embedding, learned scalar blend, spatial/temporal roles switched (V2)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(cin, cout, eps=1e-6)
        self.temporal_res_block = TemporalResnetBlock(cout, eps=1e-5)
        self.time_mixer = AlphaBlender(0.0, switch=True)

    def forward(self, x, num_frames: int):
        x = self.spatial_res_block(x)
        bf, c, h, w = x.shape
        b = bf // num_frames
        x5 = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        t = self.temporal_res_block(x5)
        out = self.time_mixer(x5, t)
        return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class Attention(nn.Module):
    """Single-head spatial self-attention of the VAE mid blocks (V4)."""

    def __init__(self, c: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        return x + o.transpose(1, 2).reshape(b, c, h, w)


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))   # V3


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Attention(c)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, boc, layers):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, co in enumerate(boc):
            self.down_blocks.append(DownEncoderBlock2D(c, co, layers, i != len(boc) - 1))
            c = co
        self.mid_block = UNetMidBlock2D(c)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class MidBlockTemporalDecoder(nn.Module):
    def __init__(self, c, layers):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(c, c) for _ in range(layers)])
        self.attentions = nn.ModuleList([Attention(c)])

    def forward(self, x, num_frames):
        x = self.resnets[0](x, num_frames)
        for res, attn in zip(self.resnets[1:], self.attentions):
            x = attn(x)
            x = res(x, num_frames)
        return x


class UpBlockTemporalDecoder(nn.Module):
    def __init__(self, cin, cout, layers, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(cin if i == 0 else cout, cout) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, num_frames):
        for r in self.resnets:
            x = r(x, num_frames)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class TemporalDecoder(nn.Module):
    def __init__(self, latent_channels, out_channels, boc, layers):
        super().__init__()
        self.conv_in = nn.Conv2d(latent_channels, boc[-1], 3, padding=1)
        self.mid_block = MidBlockTemporalDecoder(boc[-1], layers)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(UpBlockTemporalDecoder(c, co, layers + 1, i != len(boc) - 1))
            c = co
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, z, num_frames: int):
        x = self.conv_in(z)
        x = self.mid_block(x, num_frames)
        for b in self.up_blocks:
            x = b(x, num_frames)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        bf, c, h, w = x.shape
        b = bf // num_frames
        x = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        x = self.time_conv_out(x)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class _Posterior:
    def __init__(self, moments):
        self.mean, self.logvar = moments.chunk(2, dim=1)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        std = torch.exp(0.5 * self.logvar.clamp(-30.0, 20.0))
        return self.mean + std * torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)


class AutoencoderKLTemporalDecoder(nn.Module):
    """Default arguments are the SVD ``vae/config.json``."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block: int = 2, latent_channels: int = 4, scaling_factor: float = 0.18215,
                 force_upcast: bool = True):
        super().__init__()
        boc = tuple(block_out_channels)
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, latent_channels=latent_channels,
                                      scaling_factor=scaling_factor, force_upcast=force_upcast)
        self.encoder = Encoder(in_channels, latent_channels, boc, layers_per_block)
        self.decoder = TemporalDecoder(latent_channels, out_channels, boc, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def encode(self, x):
        return SimpleNamespace(latent_dist=_Posterior(self.quant_conv(self.encoder(x))))

    def decode(self, z, num_frames: int):
        return SimpleNamespace(sample=self.decoder(z, num_frames))


def tiny_vae_config(**over) -> dict:
    """Every block type at widths the kernels tile (multiples of 64), small enough for CPU fp32 tests."""
    cfg = dict(block_out_channels=(64, 128), layers_per_block=1)
    cfg.update(over)
    return cfg
