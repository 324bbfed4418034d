"""NativeUNet: the SVD ``UNetSpatioTemporalConditionModel`` forward pass on hand-written sm_100a kernels.

Drop-in for the UNet operator the reference calls at ``src/models/svd_unet.py:389-395`` (boundary B2,
SURVEY.md section 8b)::

    unet(sample=[B,F,8,H,W] fp16, timestep, encoder_hidden_states=[B,1,1024], added_time_ids=[B,3],
         return_dict=False) -> (Tensor[B,F,4,H,W],)

so ``StableVideoUNet(unet=NativeUNet(...))`` works without touching the wrapper.  Weights come from a
diffusers-layout ``state_dict`` (same keys as ``UNetSpatioTemporalConditionModel``) and are repacked once
into kernel layouts: conv filters as [Cout, taps*Cin] (tap-major K), q/k/v fused to [3C, C], GEGLU
projections interleaved [80 value | 80 gate] per 160-row tile, every N padded to 160.

Activations are channels-last matrices [M = B*F*H*W, C] end to end: the spatial token layout of the
transformer *is* the conv layout, so the reference's permutes/reshapes disappear; the temporal branch
is reached by index arithmetic (frame stride H*W rows) instead of a transposed copy.

Two orchestrations of the same launch sequence exist and must agree bit for bit (tests/test_gpu_unet.py):
  * ``orchestrator="c"`` (default): the class is a thin caller of the ``svdpp_unet_*`` handle of ``include/svdpp.h`` -
    weight packing, the ~740 launches of a forward, the activation arena and the whole denoising step
    (``svdpp_unet_step``) live in ``csrc/unet.cu``; Python passes pointers and a workspace tensor;
  * ``orchestrator="python"``: this module packs the weights with torch and issues every launch through ctypes
    (``native.py``) - kept for per-kernel timing (``native.PROFILE``) and as the cross-check of the C orchestration.
All arithmetic on activations is in ``libsvdpp.so`` either way.  There is no CPU path.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import native
from ..native import GEMM_BN, TAPS_1, TAPS_3X3, TAPS_T3, NativeError

_HALF_BN = GEMM_BN // 2

SVD_CONFIG = dict(in_channels=8, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                  down_attn=(True, True, True, False), addition_time_embed_dim=256,
                  projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
                  num_attention_heads=(5, 10, 20, 20), num_frames=25,
                  # GroupNorm eps per diffusers block class (kernel scalars).  The up blocks' value cannot be checked against
                  # the diffusers source here: get_up_block does not forward the UNet's resnet_eps=1e-5 to the SpatioTemporal
                  # up blocks, whose class default is 1e-6.  Override with config={"norm_eps": {"up": 1e-5}} if that is wrong.
                  norm_eps=dict(down_attn=1e-6, down=1e-5, mid=1e-5, up=1e-6, transformer=1e-6, out=1e-5))


def _ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _pad_rows(w: torch.Tensor, mult: int = GEMM_BN) -> torch.Tensor:
    n = w.shape[0]
    n_pad = _ceil_to(n, mult)
    if n_pad == n:
        return w.contiguous()
    out = torch.zeros((n_pad,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    out[:n] = w
    return out


def _pad_cols(w: torch.Tensor, mult: int = 64) -> torch.Tensor:
    k = w.shape[1]
    k_pad = _ceil_to(k, mult)
    if k_pad == k:
        return w.contiguous()
    out = torch.zeros((w.shape[0], k_pad), dtype=w.dtype, device=w.device)
    out[:, :k] = w
    return out


def interleave_geglu(w: torch.Tensor, b: Optional[torch.Tensor], half: int = _HALF_BN
                     ) -> Tuple[torch.Tensor, Optional[torch.Tensor], int]:
    """[2*inner, K] (value rows then gate rows, diffusers GEGLU chunk order) -> rows grouped per
    (2*half)-row tile as [half value | half gate] (half = 80 for the 160-wide kernel, 128 for the 256-wide
    CTA-pair kernel), inner padded to a multiple of half.  Returns (w, b, inner)."""
    inner = w.shape[0] // 2
    _HALF = half
    inner_pad = _ceil_to(inner, _HALF)
    tiles = inner_pad // _HALF

    def one(t: torch.Tensor) -> torch.Tensor:
        val, gate = t[:inner], t[inner:]
        tail = (inner_pad - inner,) + tuple(t.shape[1:])
        z = torch.zeros(tail, dtype=t.dtype, device=t.device)
        val, gate = torch.cat([val, z]), torch.cat([gate, z])
        val = val.reshape((tiles, _HALF) + tuple(t.shape[1:]))
        gate = gate.reshape((tiles, _HALF) + tuple(t.shape[1:]))
        return torch.cat([val, gate], dim=1).reshape((2 * inner_pad,) + tuple(t.shape[1:])).contiguous()

    return one(w), (one(b) if b is not None else None), inner


class _Lin:
    """A packed [N_pad, K_pad] weight with optional bias and its true output width."""
    __slots__ = ("w", "b", "n", "geglu", "impl")

    def __init__(self, w, b, n, geglu=False, impl=None):
        self.w, self.b, self.n, self.geglu, self.impl = w, b, n, geglu, impl


# "nearest 2x upsample, then Conv2d 3x3 pad 1" == four 2x2-tap convolutions on the low-resolution input, one per
# output parity: for output row 2h+py the three kernel rows land on input rows {h-1, h, h} (py = 0) or {h, h, h+1}
# (py = 1), so kernel rows that hit the same input row are summed (in fp32, rounded to fp16 once).  Same along W.
_SUBPIX_ROWS = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}   # parity -> ((input offset, kernel rows), ...)


def subpixel_taps(py: int, px: int) -> Tuple[Tuple[int, int, int], ...]:
    """Tap list (dw, dh, df) of the 2x2-tap convolution for output parity (py, px), K order (ih, iw, c)."""
    return tuple((dw, dh, 0) for dh, _ in _SUBPIX_ROWS[py] for dw, _ in _SUBPIX_ROWS[px])


def subpixel_weight(w: torch.Tensor, py: int, px: int) -> torch.Tensor:
    """[Co, Ci, 3, 3] -> [Co, 4*Ci] for output parity (py, px); tap-major K matching ``subpixel_taps``."""
    wf = w.float()
    parts = []
    for _, khs in _SUBPIX_ROWS[py]:
        for _, kws in _SUBPIX_ROWS[px]:
            parts.append(sum(wf[:, :, kh, kw] for kh in khs for kw in kws))
    return torch.cat(parts, dim=1).to(w.dtype)


def window_path_ok(W: int, C: int) -> bool:
    """Can the TMA shifted-window conv path tile rows of width W (see csrc/gemm_tc.cu)?"""
    return C % 64 == 0 and ((W >= 8 and 128 % W == 0) or W % 128 == 0)


class NativeUNet(nn.Module):
    def __init__(self, state_dict: Mapping[str, torch.Tensor], config: Optional[dict] = None,
                 device: torch.device | str = "cuda", gemm_impl: Optional[int] = None,
                 attn_impl: Optional[int] = None, orchestrator: Optional[str] = None,
                 attn_impl_long: Optional[int] = None):
        super().__init__()
        self.cfg = dict(SVD_CONFIG)
        if config:
            self.cfg.update(config)
        self.cfg["norm_eps"] = {**SVD_CONFIG["norm_eps"], **((config or {}).get("norm_eps") or {})}
        self.device_ = torch.device(device)
        if self.device_.type != "cuda":
            raise NativeError("NativeUNet needs a CUDA device (there is no CPU path)")
        native.load()
        # 0: one CTA per 128x160 tile; 2: CTA pairs (cta_group::2), 256x160; 3: CTA pairs with 256x256 tiles where
        # N % 256 == 0 and the one-CTA kernel elsewhere; SVDPP_GEMM_IMPL overrides the default for experiments
        self.gemm_impl = int(os.environ.get("SVDPP_GEMM_IMPL", "3")) if gemm_impl is None else gemm_impl
        # None: the two-tile FMHA (impl 2, P in TMEM) for long sequences, where its one CTA per SM amortises the
        # prologue, and the 2-CTAs-per-SM kernel (impl 0) for S < 1024; an int forces one kernel (1 = CUDA cores)
        self.attn_impl = attn_impl
        # spatial FMHA kernel for S >= 1024 (svdpp_attn_spatial_f16's impl): 4 = two query tiles per CTA with the
        # quarter-pipelined softmax (A/B in profiles/r2_*), 2 = the round-1 kernel
        self.attn_impl_long = int(os.environ.get("SVDPP_ATTN_IMPL_LONG", "7")) if attn_impl_long is None else attn_impl_long
        self.orchestrator = (orchestrator or os.environ.get("SVDPP_ORCHESTRATOR", "c")).lower()
        if self.orchestrator not in ("c", "python"):
            raise ValueError("orchestrator must be 'c' or 'python'")
        self.dtype = torch.float16
        self._sd = state_dict
        self._tensors: List[torch.Tensor] = []   # keeps packed weights alive / counted
        self._pos_cache: Dict[Tuple[str, int], torch.Tensor] = {}
        self._gn_ws: Optional[torch.Tensor] = None
        self._sms: Optional[int] = None
        self._handle: Optional[native.UNetHandle] = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_shape = None
        if self.orchestrator == "c":
            with torch.cuda.device(self.device_):
                sd = {k: v.detach().to(self.device_) for k, v in state_dict.items()}
                self._handle = native.UNetHandle(self.cfg, sd, gemm_impl=self.gemm_impl, attn_impl=self.attn_impl,
                                                 attn_impl_long=self.attn_impl_long)
                del sd
        else:
            native.splitk_workspace(self.device_)    # scratch of the split-K tail: before any graph capture
            self._build()
        self._sd = None

    # ------------------------------------------------------------------ C orchestration (svdpp_unet_*)
    def _workspace(self, B: int, F: int, H: int, W: int) -> torch.Tensor:
        """Activation workspace of the handle; grown (never inside a graph capture: the wrapper's first call per shape
        is eager) to svdpp_unet_workspace_bytes of the largest shape seen."""
        key = (B, F, H, W)
        if self._ws_shape is None or key not in self._ws_shape:
            need = self._handle.workspace_bytes(B, F, H, W)
            if self._ws is None or self._ws.numel() < need:
                if torch.cuda.is_current_stream_capturing():
                    raise NativeError("the UNet workspace must be sized by an eager call before CUDA-graph capture")
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device_)
            self._ws_shape = (self._ws_shape or set()) | {key}
        return self._ws

    def step_native(self, out, latent, image_latents, uncond_image_latents, enc, ids, gs, *, timestep, in_div, c_v, c_x,
                    sigma, dt, handoff=None) -> torch.Tensor:
        """One whole denoising step in one C call (``svdpp_unet_step``; reference svd_unet.py:351-439)."""
        B, _, F, H, W = latent.shape
        ws = self._workspace(B, F, H, W)
        self._handle.step(out, latent, image_latents, uncond_image_latents, enc.contiguous(), ids.contiguous(), gs, ws,
                          timestep=timestep, in_div=in_div, c_v=c_v, c_x=c_x, sigma=sigma, dt=dt, handoff=handoff)
        return out

    # ------------------------------------------------------------------ weight packing
    def _g(self, key: str) -> torch.Tensor:
        return self._sd[key].detach().to(self.device_, torch.float16)

    def _keep(self, t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if t is not None:
            self._tensors.append(t)
        return t

    def _lin(self, prefix: str, bias: bool = True) -> _Lin:
        w = self._g(prefix + ".weight")
        n = w.shape[0]
        b = self._g(prefix + ".bias") if bias else None
        return _Lin(self._keep(_pad_cols(_pad_rows(w))), self._keep(_pad_rows(b) if b is not None else None), n)

    def _conv3x3(self, prefix: str) -> _Lin:
        w = self._g(prefix + ".weight")                       # [Co, Ci, 3, 3]
        n = w.shape[0]
        w = w.permute(0, 2, 3, 1).reshape(n, -1)              # K order (kh, kw, ci)
        return _Lin(self._keep(_pad_cols(_pad_rows(w))), self._keep(_pad_rows(self._g(prefix + ".bias"))), n)

    def _conv_up(self, prefix: str):
        """Upsampler conv as four parity convolutions: [[_Lin for px in 0, 1] for py in 0, 1]."""
        w = self._g(prefix + ".weight")
        b = self._keep(_pad_rows(self._g(prefix + ".bias")))
        return [[_Lin(self._keep(_pad_cols(_pad_rows(subpixel_weight(w, py, px)))), b, w.shape[0]) for px in (0, 1)]
                for py in (0, 1)]

    def _conv_t3(self, prefix: str) -> _Lin:
        w = self._g(prefix + ".weight")[:, :, :, 0, 0]        # [Co, Ci, 3]
        n = w.shape[0]
        w = w.permute(0, 2, 1).reshape(n, -1)                 # K order (kt, ci)
        return _Lin(self._keep(_pad_rows(w)), self._keep(_pad_rows(self._g(prefix + ".bias"))), n)

    def _conv1x1(self, prefix: str) -> _Lin:
        w = self._g(prefix + ".weight")[:, :, 0, 0]
        return _Lin(self._keep(_pad_rows(w)), self._keep(_pad_rows(self._g(prefix + ".bias"))), w.shape[0])

    def _geglu(self, prefix: str) -> _Lin:
        w, b = self._g(prefix + ".weight"), self._g(prefix + ".bias")
        if self.gemm_impl == 3 and (w.shape[0] // 2) % 128 == 0:      # packed for the 256-wide pair kernel
            w, b, inner = interleave_geglu(w, b, half=128)
            return _Lin(self._keep(w), self._keep(b), inner, geglu=True, impl=3)
        w, b, inner = interleave_geglu(w, b)
        return _Lin(self._keep(w), self._keep(b), inner, geglu=True, impl=0 if self.gemm_impl == 3 else None)

    def _norm(self, prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._keep(self._g(prefix + ".weight").contiguous()), self._keep(self._g(prefix + ".bias").contiguous())

    def _small(self, prefix: str, bias: bool = True):
        return (self._keep(self._g(prefix + ".weight").contiguous()),
                self._keep(self._g(prefix + ".bias").contiguous()) if bias else None)

    def _alpha(self, key: str) -> float:
        # AlphaBlender, image_only_indicator == 0: alpha = sigmoid(mix_factor), rounded to fp16 as torch does
        a = torch.sigmoid(self._sd[key].detach().float().reshape(-1)[0]).to(torch.float16)
        return float(a)

    def _build_resblock(self, prefix: str, eps: float) -> dict:
        sp, tp = prefix + ".spatial_res_block", prefix + ".temporal_res_block"
        P = dict(eps=eps)
        P["norm1"], P["conv1"] = self._norm(sp + ".norm1"), self._conv3x3(sp + ".conv1")
        P["norm2"], P["conv2"] = self._norm(sp + ".norm2"), self._conv3x3(sp + ".conv2")
        P["shortcut"] = self._conv1x1(sp + ".conv_shortcut") if (sp + ".conv_shortcut.weight") in self._sd else None
        P["tnorm1"], P["tconv1"] = self._norm(tp + ".norm1"), self._conv_t3(tp + ".conv1")
        P["tnorm2"], P["tconv2"] = self._norm(tp + ".norm2"), self._conv_t3(tp + ".conv2")
        P["alpha"] = self._alpha(prefix + ".time_mixer.mix_factor")
        cout = P["conv1"].n
        # time_emb_proj of both branches go into one concatenated table (computed once per forward)
        P["temb_sp"] = self._reg_temb(sp + ".time_emb_proj", cout)
        P["temb_t"] = self._reg_temb(tp + ".time_emb_proj", cout)
        return P

    def _reg_temb(self, prefix: str, cout: int) -> Tuple[int, int]:
        off = self._temb_off
        self._temb_w.append(self._g(prefix + ".weight"))
        self._temb_b.append(self._g(prefix + ".bias"))
        self._temb_off += cout
        return off, cout

    def _reg_cross(self, prefix: str) -> Tuple[int, int]:
        """Cross-attention to the single CLIP token: register to_v / to_out.0 in the network-wide tables
        (all to_v as one stacked weight, all to_out as one grouped launch); returns the column slice."""
        wv = self._g(prefix + ".to_v.weight").contiguous()
        wo, bo = self._small(prefix + ".to_out.0")
        off, c = self._ca_off, wv.shape[0]
        self._ca_wv.append(wv)
        self._ca_groups.append((wo, bo, off, off))
        self._ca_off += c
        return off, c

    def _fuse_qkv(self, prefix: str) -> _Lin:
        w = torch.cat([self._g(prefix + ".to_q.weight"), self._g(prefix + ".to_k.weight"),
                       self._g(prefix + ".to_v.weight")], dim=0)
        n = w.shape[0]
        # 3C = 960 / 1920 are not multiples of 256: padding the rows by <= 7 % buys the 256x256 CTA-pair kernel
        # (the epilogue stores only the first n columns)
        if self.gemm_impl == 3 and n % 256 and _ceil_to(n, 256) <= 1.07 * n:
            return _Lin(self._keep(_pad_rows(w, 256)), None, n)
        return _Lin(self._keep(_pad_rows(w)), None, n)

    def _build_transformer(self, prefix: str, heads: int) -> dict:
        P = dict(heads=heads, eps=self.cfg["norm_eps"]["transformer"])
        P["norm"], P["proj_in"] = self._norm(prefix + ".norm"), self._lin(prefix + ".proj_in")
        s, t = prefix + ".transformer_blocks.0", prefix + ".temporal_transformer_blocks.0"
        P["norm1"], P["qkv1"], P["out1"] = self._norm(s + ".norm1"), self._fuse_qkv(s + ".attn1"), self._lin(s + ".attn1.to_out.0")
        P["ca"] = self._reg_cross(s + ".attn2")
        P["norm3"], P["ff1"], P["ff2"] = self._norm(s + ".norm3"), self._geglu(s + ".ff.net.0.proj"), self._lin(s + ".ff.net.2")
        P["t_norm_in"], P["t_ffin1"], P["t_ffin2"] = (self._norm(t + ".norm_in"), self._geglu(t + ".ff_in.net.0.proj"),
                                                      self._lin(t + ".ff_in.net.2"))
        P["t_norm1"], P["t_qkv"], P["t_out1"] = self._norm(t + ".norm1"), self._fuse_qkv(t + ".attn1"), self._lin(t + ".attn1.to_out.0")
        P["t_ca"] = self._reg_cross(t + ".attn2")
        P["t_norm3"], P["t_ff1"], P["t_ff2"] = self._norm(t + ".norm3"), self._geglu(t + ".ff.net.0.proj"), self._lin(t + ".ff.net.2")
        P["pos1"], P["pos2"] = self._small(prefix + ".time_pos_embed.linear_1"), self._small(prefix + ".time_pos_embed.linear_2")
        P["alpha"] = self._alpha(prefix + ".time_mixer.mix_factor")
        P["proj_out"] = self._lin(prefix + ".proj_out")
        P["key"] = prefix
        return P

    def _build(self) -> None:
        cfg = self.cfg
        boc = tuple(cfg["block_out_channels"])
        heads = tuple(cfg["num_attention_heads"])
        attn = tuple(cfg["down_attn"])
        L = cfg["layers_per_block"]
        self._temb_w, self._temb_b, self._temb_off = [], [], 0
        self._ca_wv, self._ca_groups, self._ca_off = [], [], 0
        self.conv_in = self._conv3x3("conv_in")
        self.time_mlp = (self._small("time_embedding.linear_1"), self._small("time_embedding.linear_2"))
        self.add_mlp = (self._small("add_embedding.linear_1"), self._small("add_embedding.linear_2"))
        self.down = []
        for i in range(len(boc)):
            blk = dict(res=[], attn=[], down=None)
            eps = self.cfg["norm_eps"]["down_attn" if attn[i] else "down"]
            for j in range(L):
                blk["res"].append(self._build_resblock(f"down_blocks.{i}.resnets.{j}", eps))
                if attn[i]:
                    blk["attn"].append(self._build_transformer(f"down_blocks.{i}.attentions.{j}", heads[i]))
            if i != len(boc) - 1:
                blk["down"] = self._conv3x3(f"down_blocks.{i}.downsamplers.0.conv")
            self.down.append(blk)
        self.mid = dict(res=[self._build_resblock("mid_block.resnets.0", self.cfg["norm_eps"]["mid"]),
                             self._build_resblock("mid_block.resnets.1", self.cfg["norm_eps"]["mid"])],
                        attn=self._build_transformer("mid_block.attentions.0", heads[-1]))
        self.up = []
        rheads, rattn = heads[::-1], attn[::-1]
        for i in range(len(boc)):
            blk = dict(res=[], attn=[], up=None)
            for j in range(L + 1):
                blk["res"].append(self._build_resblock(f"up_blocks.{i}.resnets.{j}", self.cfg["norm_eps"]["up"]))
                if rattn[i]:
                    blk["attn"].append(self._build_transformer(f"up_blocks.{i}.attentions.{j}", rheads[i]))
            if i != len(boc) - 1:
                blk["up"] = self._conv3x3(f"up_blocks.{i}.upsamplers.0.conv")
                blk["up4"] = self._conv_up(f"up_blocks.{i}.upsamplers.0.conv")
            self.up.append(blk)
        self.norm_out = self._norm("conv_norm_out")
        self.conv_out = self._conv3x3("conv_out")
        self.temb_w = self._keep(torch.cat(self._temb_w, dim=0).contiguous())
        self.temb_b = self._keep(torch.cat(self._temb_b, dim=0).contiguous())
        self.temb_total = self._temb_off
        del self._temb_w, self._temb_b
        self.ca_wv = self._keep(torch.cat(self._ca_wv, dim=0).contiguous())     # [sum C, 1024]
        self.ca_table = self._keep(native.pack_small_groups(self._ca_groups, self.device_))
        self.ca_n_groups, self.ca_total = len(self._ca_groups), self._ca_off
        self.ca_max_n = max(g[0].shape[0] for g in self._ca_groups)
        del self._ca_wv, self._ca_groups

    def weight_bytes(self) -> int:
        if self._handle is not None:
            return self._handle.weight_bytes()
        return sum(t.numel() * t.element_size() for t in self._tensors)

    # ------------------------------------------------------------------ op helpers
    def _new(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.float16, device=self.device_)

    def _gn(self, x1, norm, *, n_img, HW, eps, silu=True, x2=None, fps=1):
        C = x1.shape[1] + (0 if x2 is None else x2.shape[1])
        need = native.groupnorm_workspace_bytes(n_img, HW)
        if self._gn_ws is None or self._gn_ws.numel() * 4 < need:
            # zero-filled: the tail holds the arrival counters of the statistics kernel
            self._gn_ws = torch.zeros((need + 3) // 4, dtype=torch.float32, device=self.device_)
        out = self._new(x1.shape[0], C)
        return native.groupnorm_silu(out, x1, norm[0], norm[1], n_img=n_img, HW=HW, eps=eps, silu=silu, x2=x2,
                                     frames_per_stat=fps, workspace=self._gn_ws)

    # relative per-SM throughput of the tile shapes on many-wave problems (measured: 256x256 CTA pairs +15-25 % over
    # 128x160; 256x320 pairs a little below 256x256 where both apply; 128x128 below 128x160)
    _TILE_SPEED = {3: 1.2, 6: 1.15, 0: 1.0, 4: 0.8}

    def _impl(self, lin: _Lin, M: int = 1 << 30) -> int:
        """Tile shape for one GEMM.  The packed layout fixes it for GEGLU; otherwise, in auto mode (3), the candidate
        with the smallest estimated time = waves over the 148 SMs x tile area / relative speed.  The wave count
        decides at small M: at M = 3600 (the 9x16 level) N = 1280 is 75 pair tiles of 256x256 - one more than the 74
        CTA pairs, i.e. two waves - but 60 tiles of 256x320: 502 -> 875 TFLOP/s on the 3x3 convs there."""
        if lin.impl is not None:
            return lin.impl
        if self.gemm_impl != 3:
            return self.gemm_impl
        N, K = lin.w.shape
        sms = native.device_info()[2] if self._sms is None else self._sms
        self._sms = sms
        mt = (M + 127) // 128
        best, best_t = 0, None
        for impl, bn, pair in ((3, 256, True), (6, 320, True), (0, 160, False), (4, 128, False)):
            if N % bn or (impl == 4 and os.environ.get("SVDPP_NO_BN128")):
                continue
            # 256x320: three rotating TMEM buffers instead of two full stages - loses below K = 960 (measured)
            if impl == 6 and (K < 960 or os.environ.get("SVDPP_NO_PAIR320")):
                continue
            work = ((mt + 1) // 2) * (N // bn) if pair else mt * (N // bn)
            slots = sms // 2 if pair else sms
            t = -(-work // slots) * 128 * bn / self._TILE_SPEED[impl]
            if best_t is None or t < best_t:
                best, best_t = impl, t
        return best

    def _linear(self, a, lin: _Lin, *, a2=None, **epi):
        n_out = lin.n
        out = self._new(a.shape[0], n_out)
        return native.gemm(out, a, lin.w, bias=lin.b, a2=a2, geglu=lin.geglu, n_store=n_out,
                           impl=self._impl(lin, a.shape[0]), **epi)

    def _conv(self, a, lin: _Lin, dims, taps, **epi):
        B, F, H, W, C = dims
        M = B * F * H * W
        out = self._new(M, lin.n)
        if window_path_ok(W, C):
            return native.gemm(out, a, lin.w, bias=lin.b, conv_dims=dims, taps=taps, n_store=lin.n,
                               impl=self._impl(lin, M), **epi)
        cols = self._new(M, len(taps) * C)
        native.im2col(cols, a, B=B, F=F, H=H, W=W, Cc=C, Ho=H, Wo=W, stride=1, taps=taps)
        return native.gemm(out, cols, lin.w, bias=lin.b, n_store=lin.n, impl=self._impl(lin, M), **epi)

    def _small_mlp(self, x, l1, l2, x_add=None):
        h = self._new(x.shape[0], l1[0].shape[0])
        native.linear_small(h, x, l1[0], l1[1], x_add=x_add, act_out=1)
        y = self._new(x.shape[0], l2[0].shape[0])
        return native.linear_small(y, h, l2[0], l2[1])

    # ------------------------------------------------------------------ blocks
    def _resblock(self, x, skip, P, tembs, B, F, H, W):
        HW, n_img = H * W, B * F
        cout = P["conv1"].n
        cin = x.shape[1] + (0 if skip is None else skip.shape[1])
        a = self._gn(x, P["norm1"], n_img=n_img, HW=HW, eps=P["eps"], x2=skip)
        o, c = P["temb_sp"]
        h1 = self._conv(a, P["conv1"], (B, F, H, W, cin), TAPS_3X3, rowvec=tembs[:, o:o + c], rv_hw=HW, rv_div=F)
        a2 = self._gn(h1, P["norm2"], n_img=n_img, HW=HW, eps=P["eps"])
        r = x if P["shortcut"] is None else self._linear(x, P["shortcut"], a2=skip)
        xs = self._conv(a2, P["conv2"], (B, F, H, W, cout), TAPS_3X3, r1=r)
        t1 = self._gn(xs, P["tnorm1"], n_img=n_img, HW=HW, eps=P["eps"], fps=F)
        o, c = P["temb_t"]
        t2 = self._conv(t1, P["tconv1"], (B, F, H, W, cout), TAPS_T3, rowvec=tembs[:, o:o + c], rv_hw=HW, rv_div=F)
        t3 = self._gn(t2, P["tnorm2"], n_img=n_img, HW=HW, eps=P["eps"], fps=F)
        # blend: alpha*xs + (1-alpha)*(xs + h) = xs + (1-alpha)*h
        return self._conv(t3, P["tconv2"], (B, F, H, W, cout), TAPS_T3, alpha=1.0 - P["alpha"], r1=xs)

    def _pos_embed(self, P, F: int, C: int) -> torch.Tensor:
        key = (P["key"], F)
        if key not in self._pos_cache:
            s = self._new(F, C)
            native.sinusoid_embed(s, None, n_vals=F, dim=C, src_mod=F)
            self._pos_cache[key] = self._small_mlp(s, P["pos1"], P["pos2"])
        return self._pos_cache[key]

    def _cross_vecs(self, enc2d: torch.Tensor) -> torch.Tensor:
        """Cross-attention with a single context token: softmax over one key is 1, so each block adds
        to_out(to_v(ctx)) to every token (q/k projections and norm2 cannot influence the result).  All 32
        blocks' vectors in two launches: stacked to_v, then the grouped to_out."""
        h = self._new(enc2d.shape[0], self.ca_total)
        native.linear_small(h, enc2d, self.ca_wv, None)
        y = self._new(enc2d.shape[0], self.ca_total)
        return native.linear_small_grouped(y, h, self.ca_table, n_groups=self.ca_n_groups, max_n=self.ca_max_n)

    def _transformer(self, x, P, cvs, B, F, H, W):
        HW, n_img, M = H * W, B * F, x.shape[0]
        C, heads = x.shape[1], P["heads"]
        scale = 1.0 / math.sqrt(C // heads)
        a = P["alpha"]
        g = self._gn(x, P["norm"], n_img=n_img, HW=HW, eps=P["eps"], silu=False)
        h0 = self._linear(g, P["proj_in"])
        # --- spatial block
        n1 = native.layernorm(self._new(M, C), h0, *P["norm1"])
        qkv = self._linear(n1, P["qkv1"])
        att = native.attn_spatial(self._new(M, C), qkv, n_img=n_img, S=HW, heads=heads, q_off=0, k_off=C,
                                  v_off=2 * C, scale=scale,
                                  impl=(self.attn_impl_long if HW >= 1024 else 0) if self.attn_impl is None else self.attn_impl)
        o_, c_ = P["ca"]
        cv = cvs[:, o_:o_ + c_]
        h2 = self._linear(att, P["out1"], r1=h0, rowvec=cv, rv_hw=HW, rv_div=F)
        n3 = native.layernorm(self._new(M, C), h2, *P["norm3"])
        hs = self._linear(self._linear(n3, P["ff1"]), P["ff2"], r1=h2)
        # --- temporal block on (hs + frame-position embedding); token (b,f,p) stays at row (b*F+f)*HW+p
        pos = self._pos_embed(P, F, C)
        nin = native.layernorm(self._new(M, C), hs, *P["t_norm_in"], addvec=pos, add_hw=HW, add_mod=F)
        t1 = self._linear(self._linear(nin, P["t_ffin1"]), P["t_ffin2"], r1=hs, rowvec=pos, rv_hw=HW, rv_div=1,
                          rv_mod=F)
        n1t = native.layernorm(self._new(M, C), t1, *P["t_norm1"])
        qkvt = self._linear(n1t, P["t_qkv"])
        attt = native.attn_temporal(self._new(M, C), qkvt, B=B, F=F, HW=HW, heads=heads, q_off=0, k_off=C,
                                    v_off=2 * C, scale=scale)
        o_, c_ = P["t_ca"]
        cvt = cvs[:, o_:o_ + c_]
        t2 = self._linear(attt, P["t_out1"], r1=t1, rowvec=cvt, rv_hw=HW, rv_div=F)
        n3t = native.layernorm(self._new(M, C), t2, *P["t_norm3"])
        # blend fused into the last temporal GEMM: a*hs + (1-a)*(ff + t2)
        hb = self._linear(self._linear(n3t, P["t_ff1"]), P["t_ff2"], alpha=1.0 - a, r1=t2, beta1=1.0 - a, r2=hs,
                          beta2=a)
        return self._linear(hb, P["proj_out"], r1=x)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward_nhwc(self, x_in: torch.Tensor, t_dev: torch.Tensor, enc: torch.Tensor, ids: torch.Tensor,
                     B: int, F: int, H: int, W: int) -> torch.Tensor:
        """x_in: channels-last [B*F*H*W, 8]; t_dev: fp32 [B] on device; enc: [B,1,1024] or [B,1024];
        ids: [B,3].  Returns the channels-last prediction [B*F*H*W, 4]."""
        if self._handle is not None:
            out = self._new(x_in.shape[0], self.cfg["out_channels"])
            self._handle.forward(out, x_in, float(t_dev[0]) if not isinstance(t_dev, float) else t_dev,
                                 enc.reshape(B, -1).contiguous(), ids.contiguous(), self._workspace(B, F, H, W),
                                 B=B, F=F, H=H, W=W, nhwc=True)
            return out
        cfg = self.cfg
        boc = tuple(cfg["block_out_channels"])
        enc2d = enc.reshape(B, -1).contiguous()
        ids = ids.contiguous()
        # --- embeddings
        s_t = native.sinusoid_embed(self._new(B, boc[0]), t_dev, n_vals=B, dim=boc[0])
        e_t = self._small_mlp(s_t, *self.time_mlp)
        ad = cfg["addition_time_embed_dim"]
        s_a = native.sinusoid_embed(self._new(B * ids.shape[1], ad), ids.reshape(-1), n_vals=B * ids.shape[1], dim=ad)
        e_a = self._small_mlp(s_a.reshape(B, -1), *self.add_mlp)
        tembs = self._new(B, self.temb_total)   # every time_emb_proj(silu(emb)) of the network at once
        native.linear_small(tembs, e_t, self.temb_w, self.temb_b, x_add=e_a, act_in=1)
        cvs = self._cross_vecs(enc2d)
        # --- conv_in (8 channels: gather the 3x3 windows, K padded 72 -> 128)
        cin = x_in.shape[1]
        cols = self._new(x_in.shape[0], self.conv_in.w.shape[1])
        native.im2col(cols, x_in, B=B, F=F, H=H, W=W, Cc=cin, Ho=H, Wo=W, stride=1, taps=TAPS_3X3)
        x = native.gemm(self._new(x_in.shape[0], self.conv_in.n), cols, self.conv_in.w, bias=self.conv_in.b,
                        n_store=self.conv_in.n, impl=self._impl(self.conv_in, x_in.shape[0]))
        skips = [x]
        h, w = H, W
        for blk in self.down:
            for j, R in enumerate(blk["res"]):
                x = self._resblock(x, None, R, tembs, B, F, h, w)
                if blk["attn"]:
                    x = self._transformer(x, blk["attn"][j], cvs, B, F, h, w)
                skips.append(x)
            if blk["down"] is not None:
                C = x.shape[1]
                ho, wo = (h + 1) // 2, (w + 1) // 2
                lin = blk["down"]
                if window_path_ok(wo, C):
                    # Conv2d 3x3 stride 2 pad 1 as strided TMA windows (element stride 2 along W, rows 2*ho + dh)
                    x = native.gemm(self._new(B * F * ho * wo, lin.n), x, lin.w, bias=lin.b, conv_dims=(B, F, ho, wo, C),
                                    taps=TAPS_3X3, n_store=lin.n, impl=self._impl(lin, B * F * ho * wo), conv_stride=2,
                                    conv_in_hw=(h, w))
                else:
                    cols = self._new(B * F * ho * wo, 9 * C)
                    native.im2col(cols, x, B=B, F=F, H=h, W=w, Cc=C, Ho=ho, Wo=wo, stride=2, taps=TAPS_3X3)
                    x = self._linear(cols, lin)
                h, w = ho, wo
                skips.append(x)
        x = self._resblock(x, None, self.mid["res"][0], tembs, B, F, h, w)
        x = self._transformer(x, self.mid["attn"], cvs, B, F, h, w)
        x = self._resblock(x, None, self.mid["res"][1], tembs, B, F, h, w)
        for blk in self.up:
            for j, R in enumerate(blk["res"]):
                x = self._resblock(x, skips.pop(), R, tembs, B, F, h, w)
                if blk["attn"]:
                    x = self._transformer(x, blk["attn"][j], cvs, B, F, h, w)
            if blk["up"] is not None:
                C = x.shape[1]
                if window_path_ok(w, C) and not os.environ.get("SVDPP_NO_SUBPIXEL"):
                    # nearest 2x + 3x3 conv as four 2x2-tap convs on the low-resolution input (4/9 of the FLOPs, no
                    # upsampled tensor); each GEMM scatters its rows to one output parity
                    out = self._new(B * F * 4 * h * w, blk["up"].n)
                    for py in (0, 1):
                        for px in (0, 1):
                            lin = blk["up4"][py][px]
                            native.gemm(out, x, lin.w, bias=lin.b, conv_dims=(B, F, h, w, C), taps=subpixel_taps(py, px),
                                        n_store=lin.n, impl=self._impl(lin, B * F * h * w), out_up=(2, py, px))
                    h, w = 2 * h, 2 * w
                    x = out
                else:
                    up = native.upsample2x(self._new(B * F * 4 * h * w, C), x, n_img=B * F, H=h, W=w, Cc=C)
                    h, w = 2 * h, 2 * w
                    x = self._conv(up, blk["up"], (B, F, h, w, C), TAPS_3X3)
        a = self._gn(x, self.norm_out, n_img=B * F, HW=h * w, eps=self.cfg["norm_eps"]["out"])
        return self._conv(a, self.conv_out, (B, F, h, w, x.shape[1]), TAPS_3X3)

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                added_time_ids: torch.Tensor, return_dict: bool = False):
        """The diffusers operator protocol (boundary B2)."""
        if not sample.is_cuda:
            raise NativeError("NativeUNet.forward needs CUDA tensors (there is no CPU path)")
        B, F, C, H, W = sample.shape
        sample = sample.to(torch.float16).contiguous()
        if self._handle is not None:
            out = self._new(B, F, self.cfg["out_channels"], H, W)
            self._handle.forward(out, sample, float(timestep), encoder_hidden_states.to(torch.float16).reshape(B, -1).contiguous(),
                                 added_time_ids.to(torch.float16).contiguous(), self._workspace(B, F, H, W),
                                 B=B, F=F, H=H, W=W)
            return (out,)
        t_dev = torch.full((B,), float(timestep), dtype=torch.float32, device=sample.device)
        x_in = self._new(B * F * H * W, C)
        native.pack_unet_input(x_in, sample, (F * C * H * W, C * H * W, H * W), C, 1.0, None, None, 0,
                               B=B, F=F, H=H, W=W)
        v = self.forward_nhwc(x_in, t_dev, encoder_hidden_states.to(torch.float16),
                              added_time_ids.to(torch.float16), B, F, H, W)
        out = self._new(B, F, v.shape[1], H, W)
        native.nhwc_to_bfchw(out, v, B=B, F=F, Cc=v.shape[1], H=H, W=W)
        return (out,)

    # the reference wrapper probes these (svd_unet.py:139-157,175-194); attention here is already fused
    def enable_xformers_memory_efficient_attention(self, *a, **k) -> None:
        return None

    def set_attention_slice(self, *a, **k) -> None:
        return None


def flops_per_forward(cfg: dict, B: int, F: int, H: int, W: int) -> Dict[str, float]:
    """FLOPs (2*M*N*K; attention 4*S^2*C per sequence) of what NativeUNet EXECUTES.  Unlike the reference-side
    count of SURVEY.md section 8d this excludes the dead cross-attention q/k projections (1.8 %) and counts the three
    up-sampling convs at 4 instead of 9 taps (``conv_up``: they run as 2x2-tap parity convolutions), so throughput
    figures derived from it are not inflated by work that is not done."""
    boc = tuple(cfg["block_out_channels"])
    attn = tuple(cfg["down_attn"])
    L = cfg["layers_per_block"]
    out: Dict[str, float] = dict(conv3x3=0.0, conv_up=0.0, conv_t=0.0, conv1x1=0.0, linear=0.0, geglu_ff=0.0,
                                 attn_spatial=0.0, attn_temporal=0.0)

    def res(cin, cout, h, w):
        M = B * F * h * w
        out["conv3x3"] += 2.0 * M * cout * 9 * cin + 2.0 * M * cout * 9 * cout
        out["conv_t"] += 2 * (2.0 * M * cout * 3 * cout)
        if cin != cout:
            out["conv1x1"] += 2.0 * M * cout * cin

    def tr(c, h, w):
        M = B * F * h * w
        out["linear"] += 2.0 * M * c * c * 2            # proj_in, proj_out
        out["linear"] += 2 * (2.0 * M * c * 3 * c + 2.0 * M * c * c)   # qkv + out, spatial and temporal
        out["geglu_ff"] += 3 * (2.0 * M * c * 8 * c + 2.0 * M * 4 * c * c)
        out["attn_spatial"] += 4.0 * (h * w) ** 2 * c * B * F
        out["attn_temporal"] += 4.0 * F * F * c * B * h * w

    h, w = H, W
    M0 = B * F * h * w
    out["conv3x3"] += 2.0 * M0 * boc[0] * 9 * cfg["in_channels"]
    c = boc[0]
    skip_c = [c]
    for i, co in enumerate(boc):
        for j in range(L):
            res(c, co, h, w)
            c = co
            if attn[i]:
                tr(c, h, w)
            skip_c.append(c)
        if i != len(boc) - 1:
            h, w = (h + 1) // 2, (w + 1) // 2
            out["conv3x3"] += 2.0 * B * F * h * w * c * 9 * c
            skip_c.append(c)
    res(c, c, h, w)
    tr(c, h, w)
    res(c, c, h, w)
    rattn = attn[::-1]
    for i, co in enumerate(boc[::-1]):
        for j in range(L + 1):
            res(c + skip_c.pop(), co, h, w)
            c = co
            if rattn[i]:
                tr(c, h, w)
        if i != len(boc) - 1:
            h, w = 2 * h, 2 * w
            out["conv_up"] += 2.0 * B * F * h * w * c * 4 * c      # four parity convs of 2x2 taps on the h/2 x w/2 input
    out["conv3x3"] += 2.0 * B * F * h * w * cfg["out_channels"] * 9 * c
    out["total"] = sum(out.values())
    return out
