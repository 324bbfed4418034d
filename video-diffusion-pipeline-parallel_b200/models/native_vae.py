"""NativeVAE: the SVD ``AutoencoderKLTemporalDecoder`` (image -> latents, latents -> frames) on the sm_100a kernels.

Drop-in for the ``vae`` object of the reference's generation scripts (``scripts/generate_video_demo.py:119-143``:
``vae.encode(image).latent_dist.mode()``; ``:154-195``: ``vae.decode(chunk, num_frames=n).sample``;
``vae.config.scaling_factor`` / ``.force_upcast``, ``vae.dtype``) - SURVEY.md section 8(f) rank 3, the step after the
denoising loop: at 2.3 s per video of diffusion the reference's 4.9 s decode would be the bottleneck.

Same construction as ``NativeUNet``: weights come from a diffusers-layout ``state_dict`` and are repacked once (conv
filters tap-major, q/k/v fused, up-sampling convs as four pre-summed 2x2-tap parity filters, ``quant_conv`` folded into
the encoder's ``conv_out``); activations are channels-last fp16 matrices ``[images * H * W, C]`` end to end; every
contraction runs on the tcgen05 GEMM / implicit-GEMM convolution (fp32 accumulate), GroupNorm+SiLU on the fused bandwidth
kernel, and the single-head 512-wide attention of the mid blocks as GEMM (Q K^T) -> row softmax -> GEMM (P V).
``force_upcast`` is False here: the reference upcasts to fp32 because fp16 *library* convolutions overflow in this VAE's
last blocks; these kernels accumulate in fp32 and only store fp16 between layers.  There is no CPU path.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace
from typing import Dict, List, Mapping, Optional, Tuple

import torch
import torch.nn as nn

from .. import native
from ..native import TAPS_3X3, TAPS_T3, NativeError
from .native_unet import _Lin, _ceil_to, _pad_cols, _pad_rows, subpixel_taps, subpixel_weight, window_path_ok

VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                  latent_channels=4, scaling_factor=0.18215)
TAPS_DOWN = tuple((kw, kh, 0) for kh in range(3) for kw in range(3))   # F.pad(x, (0,1,0,1)) + stride-2 conv, padding 0


class _Posterior:
    """``latent_dist`` of ``vae.encode``: ``mode()`` is what the SVD pipeline uses (generate_video_demo.py:136)."""

    def __init__(self, mean: torch.Tensor, logvar: torch.Tensor):
        self.mean, self.logvar = mean, logvar

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self, generator=None) -> torch.Tensor:
        std = torch.exp(0.5 * self.logvar.float().clamp(-30.0, 20.0))
        eps = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=torch.float32)
        return (self.mean.float() + std * eps).to(self.mean.dtype)


class NativeVAE(nn.Module):
    def __init__(self, state_dict: Mapping[str, torch.Tensor], config: Optional[dict] = None,
                 device: torch.device | str = "cuda"):
        super().__init__()
        cfg = dict(VAE_CONFIG)
        if config:
            cfg.update({k: v for k, v in (config if isinstance(config, dict) else vars(config)).items() if k in cfg})
        self.cfg = cfg
        self.config = SimpleNamespace(**cfg, force_upcast=False)
        self.device_ = torch.device(device)
        if self.device_.type != "cuda":
            raise NativeError("NativeVAE needs a CUDA device (there is no CPU path)")
        native.load()
        self.dtype = torch.float16
        self._sd = state_dict
        self._tensors: List[torch.Tensor] = []
        self._gn_ws: Optional[torch.Tensor] = None
        self._zeros: Dict[Tuple[int, ...], torch.Tensor] = {}
        native.splitk_workspace(self.device_)
        self._build()
        self._sd = None

    @classmethod
    def from_pretrained(cls, model_id: str = "stabilityai/stable-video-diffusion-img2vid-xt", subfolder: str = "vae",
                        torch_dtype: torch.dtype = torch.float16, device: torch.device | str = "cuda",
                        config: Optional[dict] = None, **_) -> "NativeVAE":
        """``AutoencoderKLTemporalDecoder.from_pretrained(model_id, subfolder="vae", torch_dtype=...)`` of reference
        ``scripts/generate_video_demo.py:263-267``: ``model_id`` is a local snapshot directory (hub layout) or
        ``random-init[:seed]``; there is no network (``frontend_weights.py``)."""
        from . import frontend_weights as fw
        if torch_dtype != torch.float16:
            raise NativeError("NativeVAE stores fp16 and accumulates in fp32; torch_dtype must be torch.float16")
        if fw.is_random_init(model_id):
            sd = fw.random_state_dict(fw.vae_param_shapes(config), seed=fw.random_init_seed(model_id) + 1, device=device)
        else:
            sd, file_cfg = fw.load_component(model_id, subfolder, device=device)
            config = {**(file_cfg or {}), **(config or {})}
        return cls(sd, config=config, device=device)

    # ------------------------------------------------------------------ weight packing
    def _g(self, key: str) -> torch.Tensor:
        return self._sd[key].detach().to(self.device_, torch.float16)

    def _keep(self, t):
        if t is not None:
            self._tensors.append(t)
        return t

    @staticmethod
    def _tile(n: int) -> int:
        """Tile width the weight is padded to: 256-wide CTA pairs where N allows, else 128-wide tiles."""
        return 256 if n % 256 == 0 else 128

    def _pack(self, w2d: torch.Tensor, b: Optional[torch.Tensor]) -> _Lin:
        n = w2d.shape[0]
        t = self._tile(n)
        # N = 128 layers: one CTA per 128x128 tile (impl 4), or CTA pairs on 256x128 tiles (impl 7, SVDPP_VAE_PAIR128=1)
        impl = 3 if t == 256 else (7 if os.environ.get("SVDPP_VAE_PAIR128", "0") not in ("", "0") else 4)
        return _Lin(self._keep(_pad_cols(_pad_rows(w2d, t))), self._keep(_pad_rows(b, t)) if b is not None else None, n,
                    impl=impl)

    def _conv3x3(self, prefix: str, pad_ci: int = 0) -> _Lin:
        w = self._g(prefix + ".weight")                       # [Co, Ci, 3, 3]
        if pad_ci and w.shape[1] < pad_ci:                    # conv_in: 3 / 4 input channels padded to 8 (zeros)
            z = torch.zeros((w.shape[0], pad_ci - w.shape[1], 3, 3), dtype=w.dtype, device=w.device)
            w = torch.cat([w, z], dim=1)
        return self._pack(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), self._g(prefix + ".bias"))

    def _conv_up(self, prefix: str):
        w, b = self._g(prefix + ".weight"), self._g(prefix + ".bias")
        return [[self._pack(subpixel_weight(w, py, px), b) for px in (0, 1)] for py in (0, 1)]

    def _conv_t3(self, prefix: str) -> _Lin:
        w = self._g(prefix + ".weight")[:, :, :, 0, 0]        # [Co, Ci, 3]
        return self._pack(w.permute(0, 2, 1).reshape(w.shape[0], -1), self._g(prefix + ".bias"))

    def _conv1x1(self, prefix: str) -> _Lin:
        return self._pack(self._g(prefix + ".weight")[:, :, 0, 0], self._g(prefix + ".bias"))

    def _norm(self, prefix: str):
        return self._keep(self._g(prefix + ".weight").contiguous()), self._keep(self._g(prefix + ".bias").contiguous())

    def _resnet(self, prefix: str) -> dict:
        P = dict(norm1=self._norm(prefix + ".norm1"), conv1=self._conv3x3(prefix + ".conv1"),
                 norm2=self._norm(prefix + ".norm2"), conv2=self._conv3x3(prefix + ".conv2"))
        P["shortcut"] = self._conv1x1(prefix + ".conv_shortcut") if (prefix + ".conv_shortcut.weight") in self._sd else None
        return P

    def _st_resblock(self, prefix: str) -> dict:
        P = self._resnet(prefix + ".spatial_res_block")
        t = prefix + ".temporal_res_block"
        P.update(tnorm1=self._norm(t + ".norm1"), tconv1=self._conv_t3(t + ".conv1"), tnorm2=self._norm(t + ".norm2"),
                 tconv2=self._conv_t3(t + ".conv2"))
        # AlphaBlender "learned", switch_spatial_to_temporal_mix: out = (1 - s) * spatial + s * temporal, s = sigmoid(mix);
        # temporal = spatial + h, so out = spatial + s * h
        P["s"] = float(torch.sigmoid(self._sd[prefix + ".time_mixer.mix_factor"].detach().float().reshape(-1)[0]))
        return P

    def _attention(self, prefix: str) -> dict:
        w = torch.cat([self._g(prefix + f".to_{n}.weight") for n in "qkv"], dim=0)
        b = torch.cat([self._g(prefix + f".to_{n}.bias") for n in "qkv"], dim=0)
        return dict(norm=self._norm(prefix + ".group_norm"), qkv=self._pack(w, b),
                    out=self._pack(self._g(prefix + ".to_out.0.weight"), self._g(prefix + ".to_out.0.bias")),
                    c=w.shape[1])

    def _build(self) -> None:
        boc = tuple(self.cfg["block_out_channels"])
        L = self.cfg["layers_per_block"]
        lat = self.cfg["latent_channels"]
        # ---- encoder (standard AutoencoderKL encoder)
        E = dict(conv_in=self._conv3x3("encoder.conv_in", pad_ci=8), down=[])
        for i in range(len(boc)):
            blk = dict(res=[self._resnet(f"encoder.down_blocks.{i}.resnets.{j}") for j in range(L)], down=None)
            if i != len(boc) - 1:
                blk["down"] = self._conv3x3(f"encoder.down_blocks.{i}.downsamplers.0.conv")
            E["down"].append(blk)
        E["mid"] = dict(res=[self._resnet("encoder.mid_block.resnets.0"), self._resnet("encoder.mid_block.resnets.1")],
                        attn=self._attention("encoder.mid_block.attentions.0"))
        E["norm_out"] = self._norm("encoder.conv_norm_out")
        # quant_conv (1x1, 8 -> 8) folded into conv_out: Wq (Wc * h + bc) + bq, composed in fp32
        wc = self._sd["encoder.conv_out.weight"].detach().to(self.device_, torch.float32)        # [2L, C, 3, 3]
        bc = self._sd["encoder.conv_out.bias"].detach().to(self.device_, torch.float32)
        wq = self._sd["quant_conv.weight"].detach().to(self.device_, torch.float32)[:, :, 0, 0]  # [2L, 2L]
        bq = self._sd["quant_conv.bias"].detach().to(self.device_, torch.float32)
        wf = torch.einsum("om,mchw->ochw", wq, wc)
        E["conv_out"] = self._pack(wf.permute(0, 2, 3, 1).reshape(wf.shape[0], -1).half(), (wq @ bc + bq).half())
        self.enc = E
        # ---- temporal decoder
        D = dict(conv_in=self._conv3x3("decoder.conv_in", pad_ci=8))
        D["mid"] = dict(res=[self._st_resblock(f"decoder.mid_block.resnets.{j}") for j in range(L)],
                        attn=self._attention("decoder.mid_block.attentions.0"))
        D["up"] = []
        for i in range(len(boc)):
            blk = dict(res=[self._st_resblock(f"decoder.up_blocks.{i}.resnets.{j}") for j in range(L + 1)], up4=None)
            if i != len(boc) - 1:
                blk["up4"] = self._conv_up(f"decoder.up_blocks.{i}.upsamplers.0.conv")
                blk["up"] = self._conv3x3(f"decoder.up_blocks.{i}.upsamplers.0.conv")
            D["up"].append(blk)
        D["norm_out"] = self._norm("decoder.conv_norm_out")
        # 3 output channels stored 4 wide (a zero filter): rows of the [M, 4] result are 8-byte aligned, as the UNet's conv_out
        wo, bo = self._g("decoder.conv_out.weight"), self._g("decoder.conv_out.bias")
        wo = torch.cat([wo, torch.zeros((1,) + tuple(wo.shape[1:]), dtype=wo.dtype, device=wo.device)], dim=0)
        bo = torch.cat([bo, torch.zeros(1, dtype=bo.dtype, device=bo.device)])
        D["conv_out"] = self._pack(wo.permute(0, 2, 3, 1).reshape(wo.shape[0], -1), bo)
        D["time_w"] = self._keep(self._g("decoder.time_conv_out.weight")[:, :, :, 0, 0].contiguous())    # [3, 3, 3] (co, ci, kt)
        D["time_b"] = self._keep(self._g("decoder.time_conv_out.bias").contiguous())
        self.dec = D
        self.latent_channels = lat

    def weight_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._tensors)

    # ------------------------------------------------------------------ op helpers
    def _new(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.float16, device=self.device_)

    def _gn(self, x, norm, *, n_img, HW, eps, silu=True, fps=1):
        need = native.groupnorm_workspace_bytes(n_img, HW)
        if self._gn_ws is None or self._gn_ws.numel() * 4 < need:
            self._gn_ws = torch.zeros((need + 3) // 4, dtype=torch.float32, device=self.device_)
        return native.groupnorm_silu(self._new(x.shape[0], x.shape[1]), x, norm[0], norm[1], n_img=n_img, HW=HW, eps=eps,
                                     silu=silu, frames_per_stat=fps, workspace=self._gn_ws)

    def _linear(self, a, lin: _Lin, **epi):
        return native.gemm(self._new(a.shape[0], lin.n), a, lin.w, bias=lin.b, n_store=lin.n, impl=lin.impl, **epi)

    def _conv(self, a, lin: _Lin, dims, taps, **epi):
        B, F, H, W, C = dims
        M = B * F * H * W
        out = self._new(M, lin.n)
        if window_path_ok(W, C):
            return native.gemm(out, a, lin.w, bias=lin.b, conv_dims=dims, taps=taps, n_store=lin.n, impl=lin.impl, **epi)
        cols = self._new(M, lin.w.shape[1])
        native.im2col(cols, a, B=B, F=F, H=H, W=W, Cc=C, Ho=H, Wo=W, stride=1, taps=taps)
        return native.gemm(out, cols, lin.w, bias=lin.b, n_store=lin.n, impl=lin.impl, **epi)

    def _to_nhwc8(self, x: torch.Tensor, div: float = 1.0) -> torch.Tensor:
        """[N, C <= 8, H, W] (any float dtype) -> channels-last [N*H*W, 8] fp16, x / div in the first C channels."""
        N, C, H, W = x.shape
        x = x.to(torch.float16).contiguous()
        key = (N, 8 - C, H, W)
        if key not in self._zeros:
            self._zeros[key] = torch.zeros(key, dtype=torch.float16, device=self.device_)
        out = self._new(N * H * W, 8)
        native.pack_unet_input(out, x, (C * H * W, 0, H * W), C, div, self._zeros[key], ((8 - C) * H * W, 0, H * W), 8 - C,
                               B=N, F=1, H=H, W=W)
        return out

    # ------------------------------------------------------------------ blocks
    def _resnet_fwd(self, x, P, n_img, H, W, eps=1e-6):
        HW = H * W
        cin, cout = x.shape[1], P["conv1"].n
        a = self._gn(x, P["norm1"], n_img=n_img, HW=HW, eps=eps)
        h = self._conv(a, P["conv1"], (n_img, 1, H, W, cin), TAPS_3X3)
        a2 = self._gn(h, P["norm2"], n_img=n_img, HW=HW, eps=eps)
        r = x if P["shortcut"] is None else self._linear(x, P["shortcut"])
        return self._conv(a2, P["conv2"], (n_img, 1, H, W, cout), TAPS_3X3, r1=r)

    def _st_resblock_fwd(self, x, P, B, F, H, W):
        HW, n_img = H * W, B * F
        xs = self._resnet_fwd(x, P, n_img, H, W, eps=1e-6)
        c = xs.shape[1]
        t1 = self._gn(xs, P["tnorm1"], n_img=n_img, HW=HW, eps=1e-5, fps=F)
        t2 = self._conv(t1, P["tconv1"], (B, F, H, W, c), TAPS_T3)
        t3 = self._gn(t2, P["tnorm2"], n_img=n_img, HW=HW, eps=1e-5, fps=F)
        return self._conv(t3, P["tconv2"], (B, F, H, W, c), TAPS_T3, alpha=P["s"], r1=xs)

    def _attention_fwd(self, x, P, n_img, H, W):
        """Single head of width C: softmax(Q K^T / sqrt(C)) V per image, as GEMM -> row softmax -> GEMM."""
        S, C = H * W, P["c"]
        if S % 128:
            raise NativeError(f"VAE attention needs H*W to be a multiple of 128 (got {H}x{W})")
        g = self._gn(x, P["norm"], n_img=n_img, HW=S, eps=1e-6, silu=False)
        qkv = self._linear(g, P["qkv"])
        o = self._new(n_img * S, C)
        scores, vt = self._new(S, S), self._new(C, S)
        s_impl, o_impl = (3 if S % 256 == 0 else 4), (3 if C % 256 == 0 else 4)
        for n in range(n_img):
            rows = slice(n * S, (n + 1) * S)
            q, k, v = qkv[rows, 0:C], qkv[rows, C:2 * C], qkv[rows, 2 * C:3 * C]
            native.gemm(scores, q, k, n_store=S, impl=s_impl)
            native.softmax_rows(scores, 1.0 / math.sqrt(C))
            native.transpose(vt, v)
            native.gemm(o[rows], scores, vt, n_store=C, impl=o_impl)
        return self._linear(o, P["out"], r1=x)

    # ------------------------------------------------------------------ encode / decode
    @torch.no_grad()
    def encode(self, x: torch.Tensor):
        """``x``: [N, 3, H, W] in [-1, 1].  Returns an object with ``.latent_dist.mode()`` -> [N, 4, H/8, W/8]."""
        if not x.is_cuda:
            raise NativeError("NativeVAE.encode needs CUDA tensors (there is no CPU path)")
        N, _, H, W = x.shape
        E = self.enc
        h = self._conv(self._to_nhwc8(x), E["conv_in"], (N, 1, H, W, 8), TAPS_3X3)
        for blk in E["down"]:
            for R in blk["res"]:
                h = self._resnet_fwd(h, R, N, H, W)
            if blk["down"] is not None:
                C, Ho, Wo = h.shape[1], H // 2, W // 2
                lin = blk["down"]
                if window_path_ok(Wo, C):
                    h = native.gemm(self._new(N * Ho * Wo, lin.n), h, lin.w, bias=lin.b, conv_dims=(N, 1, Ho, Wo, C),
                                    taps=TAPS_DOWN, n_store=lin.n, impl=lin.impl, conv_stride=2, conv_in_hw=(H, W))
                else:
                    cols = self._new(N * Ho * Wo, lin.w.shape[1])
                    native.im2col(cols, h, B=N, F=1, H=H, W=W, Cc=C, Ho=Ho, Wo=Wo, stride=2, taps=TAPS_DOWN)
                    h = native.gemm(self._new(N * Ho * Wo, lin.n), cols, lin.w, bias=lin.b, n_store=lin.n, impl=lin.impl)
                H, W = Ho, Wo
        h = self._resnet_fwd(h, E["mid"]["res"][0], N, H, W)
        h = self._attention_fwd(h, E["mid"]["attn"], N, H, W)
        h = self._resnet_fwd(h, E["mid"]["res"][1], N, H, W)
        a = self._gn(h, E["norm_out"], n_img=N, HW=H * W, eps=1e-6)
        m = self._conv(a, E["conv_out"], (N, 1, H, W, a.shape[1]), TAPS_3X3)            # moments (quant_conv folded in)
        L = self.latent_channels
        mom = self._new(N, 1, 2 * L, H, W)
        native.nhwc_to_bfchw(mom, m, B=N, F=1, Cc=2 * L, H=H, W=W)
        mom = mom.reshape(N, 2 * L, H, W)
        return SimpleNamespace(latent_dist=_Posterior(mom[:, :L].contiguous(), mom[:, L:].contiguous()))

    @torch.no_grad()
    def decode(self, z: torch.Tensor, num_frames: int = 1, out_dtype: torch.dtype = torch.float16):
        """``z``: [B*F, 4, h, w] (already divided by ``scaling_factor``, as generate_video_demo.py:168 does).
        Returns an object with ``.sample`` = frames [B*F, 3, 8h, 8w]."""
        if not z.is_cuda:
            raise NativeError("NativeVAE.decode needs CUDA tensors (there is no CPU path)")
        BF, _, H, W = z.shape
        F = num_frames
        if BF % F:
            raise ValueError("the number of latents must be a multiple of num_frames")
        B = BF // F
        D = self.dec
        x = self._conv(self._to_nhwc8(z), D["conv_in"], (BF, 1, H, W, 8), TAPS_3X3)
        x = self._st_resblock_fwd(x, D["mid"]["res"][0], B, F, H, W)
        for R in D["mid"]["res"][1:]:
            x = self._attention_fwd(x, D["mid"]["attn"], BF, H, W)
            x = self._st_resblock_fwd(x, R, B, F, H, W)
        for blk in D["up"]:
            for R in blk["res"]:
                x = self._st_resblock_fwd(x, R, B, F, H, W)
            if blk["up4"] is not None:
                C = x.shape[1]
                if window_path_ok(W, C):
                    out = self._new(BF * 4 * H * W, blk["up"].n)
                    for py in (0, 1):
                        for px in (0, 1):
                            lin = blk["up4"][py][px]
                            native.gemm(out, x, lin.w, bias=lin.b, conv_dims=(BF, 1, H, W, C), taps=subpixel_taps(py, px),
                                        n_store=lin.n, impl=lin.impl, out_up=(2, py, px))
                    x = out
                    H, W = 2 * H, 2 * W
                else:
                    up = native.upsample2x(self._new(BF * 4 * H * W, C), x, n_img=BF, H=H, W=W, Cc=C)
                    H, W = 2 * H, 2 * W
                    x = self._conv(up, blk["up"], (BF, 1, H, W, C), TAPS_3X3)
        a = self._gn(x, D["norm_out"], n_img=BF, HW=H * W, eps=1e-6)
        del x
        y = self._conv(a, D["conv_out"], (BF, 1, H, W, a.shape[1]), TAPS_3X3)           # [M, 4] channels-last (3 used)
        del a
        out = torch.empty((BF, 3, H, W), dtype=out_dtype, device=self.device_)
        native.time_conv_out(out, y, D["time_w"], D["time_b"], B=B, F=F, HW=H * W)
        return SimpleNamespace(sample=out)

    def to(self, *args, **kwargs):  # the reference flips the VAE between fp16 and fp32 (force_upcast); nothing to do here
        return self
