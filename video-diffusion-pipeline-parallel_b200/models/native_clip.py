"""NativeCLIPVision: the CLIP image encoder of Stable Video Diffusion (``CLIPVisionModelWithProjection``, ViT-H/14:
hidden 1280, 32 layers, 16 heads of width 80, 257 tokens, projection 1024) on the sm_100a kernels.

Drop-in for the ``image_encoder`` of the reference's generation script (``scripts/generate_video_demo.py:112-117``:
``image_encoder(pixel_values).image_embeds`` -> ``[B, 1024]``) - SURVEY.md section 8(f) rank 3.  Weights come from a
transformers-layout ``state_dict`` (``vision_model.encoder.layers.0.self_attn.q_proj.weight`` ...), so the oracle for
this module is the real ``transformers`` class with the same weights (tests/kernel_checks.py::clip_vision) - parity here
is pinned against the library itself, not against a restatement.

How the shapes meet the kernels (head width 80 and 257 tokens are not tile multiples):
  * every image's 257 tokens live in 384 rows (3 x 128); the 127 padding rows stay finite and are never read back;
  * q / k / v / out projections are packed with each head's 80 channels padded to 128 (zero rows / columns); attention
    of all (image, head) pairs of a layer is ONE launch of ``svdpp_attn_small_f16`` (K / V of a head resident in shared
    memory, fp32 maths; it writes zeros into the padding columns and padding-token rows).  ``attn="gemm"`` keeps the
    first implementation - per head GEMM (Q K^T, 384 x 384, K = 128) -> row softmax with the 127 padding keys masked
    (``svdpp_softmax_rows(n_valid=257)``) -> transpose -> GEMM (P V): 2048 launches per image, 27 ms against 10 ms for
    the library, which is why the one-launch kernel exists;
  * the patch embedding (14 x 14 stride 14 convolution, no bias) is a GEMM over unfolded patches (K = 588 padded to 640)
    whose epilogue adds the position embedding row (``rowvec``) and writes straight into the token matrix;
  * ``fc1`` + exact GELU is the GEGLU epilogue with a constant value branch (zero weights, bias 1: 1 * gelu(gate)).
LayerNorms use the fused bandwidth kernel.  There is no CPU path.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Mapping, Optional

import torch
import torch.nn as nn

from .. import native
from ..native import NativeError
from .native_unet import _Lin, _pad_cols, _pad_rows, interleave_geglu

CLIP_VIT_H = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224,
                  patch_size=14, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5)
HD_PAD = 128      # every head's channels padded to one 128-wide K block
S_TILE = 128      # tokens per image padded to a multiple of this


class NativeCLIPVision(nn.Module):
    def __init__(self, state_dict: Mapping[str, torch.Tensor], config: Optional[dict] = None,
                 device: torch.device | str = "cuda", attn: str = "fused", use_graph: bool = False):
        super().__init__()
        self.use_graph = bool(use_graph)      # replay the launch sequence of a batch size as one CUDA graph
        self._graphs = {}
        if attn not in ("fused", "gemm"):
            raise ValueError(f"attn must be 'fused' or 'gemm', got {attn!r}")
        self.attn = attn
        cfg = dict(CLIP_VIT_H)
        if config is not None:
            src = config if isinstance(config, dict) else config.to_dict()
            cfg.update({k: src[k] for k in cfg if k in src})
        if cfg["hidden_act"] != "gelu":
            raise NativeError(f"hidden_act {cfg['hidden_act']!r}: only the exact-erf GELU of the SVD image encoder is built")
        self.cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self.device_ = torch.device(device)
        if self.device_.type != "cuda":
            raise NativeError("NativeCLIPVision needs a CUDA device (there is no CPU path)")
        native.load()
        self.dtype = torch.float16
        C, H = cfg["hidden_size"], cfg["num_attention_heads"]
        self.hd = C // H
        if self.hd > HD_PAD or C % 8:
            raise NativeError("head width must be <= 128")
        self.n_patches = (cfg["image_size"] // cfg["patch_size"]) ** 2
        self.S = self.n_patches + 1
        self.S_pad = (self.S + S_TILE - 1) // S_TILE * S_TILE
        self._sd = state_dict
        self._tensors: List[torch.Tensor] = []
        self._build()
        self._sd = None

    @classmethod
    def from_pretrained(cls, model_id: str = "stabilityai/stable-video-diffusion-img2vid-xt",
                        subfolder: str = "image_encoder", torch_dtype: torch.dtype = torch.float16,
                        device: torch.device | str = "cuda", config: Optional[dict] = None, **kwargs) -> "NativeCLIPVision":
        """``CLIPVisionModelWithProjection.from_pretrained(model_id, subfolder="image_encoder", torch_dtype=...)`` of
        reference ``scripts/generate_video_demo.py:252-255``: a local snapshot directory (hub layout) or
        ``random-init[:seed]``; there is no network (``frontend_weights.py``)."""
        from . import frontend_weights as fw
        if torch_dtype != torch.float16:
            raise NativeError("NativeCLIPVision computes in fp16 with fp32 accumulation; torch_dtype must be torch.float16")
        if fw.is_random_init(model_id):
            sd = fw.random_state_dict(fw.clip_param_shapes(config), seed=fw.random_init_seed(model_id) + 2, device=device)
        else:
            sd, file_cfg = fw.load_component(model_id, subfolder, device=device)
            config = {**(file_cfg or {}), **(config or {})}
        return cls(sd, config=config, device=device, **kwargs)

    # ------------------------------------------------------------------ weight packing
    def _g(self, key: str) -> torch.Tensor:
        return self._sd[key].detach().to(self.device_, torch.float16)

    def _keep(self, t):
        if t is not None:
            self._tensors.append(t)
        return t

    def _pack(self, w: torch.Tensor, b: Optional[torch.Tensor]) -> _Lin:
        n = w.shape[0]
        t = 256 if n % 256 == 0 else 128
        return _Lin(self._keep(_pad_cols(_pad_rows(w, t))), self._keep(_pad_rows(b, t)) if b is not None else None, n,
                    impl=3 if t == 256 else 4)

    def _head_pad_rows(self, w: torch.Tensor, b: torch.Tensor):
        """[heads*hd, K] -> [heads*128, K]: each head's rows followed by zero rows (same for the bias)."""
        H, hd = self.cfg["num_attention_heads"], self.hd
        wp = torch.zeros((H, HD_PAD, w.shape[1]), dtype=w.dtype, device=w.device)
        wp[:, :hd] = w.reshape(H, hd, -1)
        bp = torch.zeros((H, HD_PAD), dtype=b.dtype, device=b.device)
        bp[:, :hd] = b.reshape(H, hd)
        return wp.reshape(H * HD_PAD, -1), bp.reshape(-1)

    def _norm(self, prefix: str):
        return self._keep(self._g(prefix + ".weight").contiguous()), self._keep(self._g(prefix + ".bias").contiguous())

    def _build(self) -> None:
        cfg = self.cfg
        C, H, hd = cfg["hidden_size"], cfg["num_attention_heads"], self.hd
        e = "vision_model.embeddings."
        wp = self._g(e + "patch_embedding.weight")                      # [C, 3, p, p] -> [C, 3*p*p], K order (c, kh, kw)
        self.patch = self._pack(wp.reshape(C, -1), None)
        pos = self._g(e + "position_embedding.weight")                  # [S, C]
        self.pos_patches = self._keep(pos[1:].contiguous())             # added to patch i by the GEMM epilogue
        self.cls_row = self._keep((self._g(e + "class_embedding") + pos[0]).contiguous())     # fp16 add, as the library
        self.pre_ln = self._norm("vision_model.pre_layrnorm")
        self.layers = []
        for i in range(cfg["num_hidden_layers"]):
            p = f"vision_model.encoder.layers.{i}."
            ws, bs = [], []
            for n in ("q_proj", "k_proj", "v_proj"):
                w_, b_ = self._head_pad_rows(self._g(p + f"self_attn.{n}.weight"), self._g(p + f"self_attn.{n}.bias"))
                ws.append(w_)
                bs.append(b_)
            wo = self._g(p + "self_attn.out_proj.weight")               # [C, heads*hd] -> columns padded per head
            wop = torch.zeros((C, H, HD_PAD), dtype=wo.dtype, device=wo.device)
            wop[:, :, :hd] = wo.reshape(C, H, hd)
            w1, b1 = self._g(p + "mlp.fc1.weight"), self._g(p + "mlp.fc1.bias")
            # exact GELU through the GEGLU epilogue: value branch = 0 * x + 1, gate branch = fc1
            wg, bg, inner = interleave_geglu(torch.cat([torch.zeros_like(w1), w1]), torch.cat([torch.ones_like(b1), b1]),
                                             half=128)
            self.layers.append(dict(
                ln1=self._norm(p + "layer_norm1"), qkv=self._pack(torch.cat(ws), torch.cat(bs)),
                out=self._pack(wop.reshape(C, H * HD_PAD), self._g(p + "self_attn.out_proj.bias")),
                ln2=self._norm(p + "layer_norm2"), fc1=_Lin(self._keep(wg), self._keep(bg), inner, geglu=True, impl=3),
                fc2=self._pack(self._g(p + "mlp.fc2.weight"), self._g(p + "mlp.fc2.bias"))))
        self.post_ln = self._norm("vision_model.post_layernorm")
        self.proj = self._keep(self._g("visual_projection.weight").contiguous())                # [P, C], no bias

    def weight_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._tensors)

    # ------------------------------------------------------------------ forward
    def _new(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.float16, device=self.device_)

    def _linear(self, a, lin: _Lin, **epi):
        return native.gemm(self._new(a.shape[0], lin.n), a, lin.w, bias=lin.b, geglu=lin.geglu, n_store=lin.n, impl=lin.impl,
                           **epi)

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, **_):
        """``pixel_values``: [B, 3, image_size, image_size] (the feature extractor's output).  Returns an object with
        ``.image_embeds`` [B, projection_dim] (and ``.last_hidden_state`` [B, 257, hidden]).  With ``use_graph`` the
        ~230 launches of a batch size are captured once (after one eager run) and replayed; the returned tensors are then
        the graph's static outputs, overwritten by the next call."""
        if not pixel_values.is_cuda:
            raise NativeError("NativeCLIPVision needs CUDA tensors (there is no CPU path)")
        if not self.use_graph:
            return self._forward(pixel_values)
        key = tuple(pixel_values.shape)
        ent = self._graphs.get(key)
        if ent is None:
            static_in = pixel_values.to(torch.float16).clone()
            self._forward(static_in)                       # eager once: function attributes, tensor-map cache, allocator
            torch.cuda.synchronize(self.device_)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._forward(static_in)
            ent = self._graphs[key] = (g, static_in, static_out)
        g, static_in, static_out = ent
        static_in.copy_(pixel_values)
        g.replay()
        return static_out

    def _forward(self, pixel_values: torch.Tensor):
        cfg = self.cfg
        B, _, Hi, Wi = pixel_values.shape
        ps, C, H = cfg["patch_size"], cfg["hidden_size"], cfg["num_attention_heads"]
        if Hi != cfg["image_size"] or Wi != cfg["image_size"]:
            raise ValueError(f"Input image size ({Hi}*{Wi}) doesn't match model ({cfg['image_size']}*{cfg['image_size']}).")
        g, S, Sp, NP = Hi // ps, self.S, self.S_pad, self.n_patches
        # unfold the non-overlapping patches (pure data movement): [B, 3, g, p, g, p] -> [B*g*g, 3*p*p], K padded to 64
        x = pixel_values.to(torch.float16).reshape(B, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(B * NP, 3 * ps * ps)
        xp = torch.zeros((B * NP, self.patch.w.shape[1]), dtype=torch.float16, device=self.device_)
        xp[:, :x.shape[1]] = x
        tok = torch.zeros((B * Sp, C), dtype=torch.float16, device=self.device_)
        for b in range(B):      # patch embedding + position embedding, written into rows 1..256 of the image's block
            native.gemm(tok[b * Sp + 1: b * Sp + 1 + NP], xp[b * NP:(b + 1) * NP], self.patch.w, n_store=C, impl=self.patch.impl,
                        rowvec=self.pos_patches, rv_hw=1, rv_div=1, rv_mod=NP)
            tok[b * Sp] = self.cls_row
        h = native.layernorm(self._new(B * Sp, C), tok, *self.pre_ln, eps=cfg["layer_norm_eps"])
        scale = 1.0 / math.sqrt(self.hd)
        scores, vt = self._new(Sp, Sp), self._new(HD_PAD, Sp)
        s_impl = 3 if Sp % 256 == 0 else 4
        for L in self.layers:
            n1 = native.layernorm(self._new(B * Sp, C), h, *L["ln1"], eps=cfg["layer_norm_eps"])
            qkv = self._linear(n1, L["qkv"])                                  # [B*Sp, 3 * heads * 128]
            o = self._new(B * Sp, H * HD_PAD)
            if self.attn == "fused":
                native.attn_small(o, qkv, n_img=B, S=S, S_pad=Sp, heads=H, head_dim=self.hd, q_off=0, k_off=H * HD_PAD,
                                  v_off=2 * H * HD_PAD, head_stride=HD_PAD, out_head_stride=HD_PAD, scale=scale)
            for b in range(B if self.attn == "gemm" else 0):
                rows = slice(b * Sp, (b + 1) * Sp)
                for hh in range(H):
                    q = qkv[rows, hh * HD_PAD:(hh + 1) * HD_PAD]
                    k = qkv[rows, (H + hh) * HD_PAD:(H + hh + 1) * HD_PAD]
                    v = qkv[rows, (2 * H + hh) * HD_PAD:(2 * H + hh + 1) * HD_PAD]
                    native.gemm(scores, q, k, n_store=Sp, impl=s_impl)
                    native.softmax_rows(scores, scale, n_valid=S)
                    native.transpose(vt, v)
                    native.gemm(o[rows, hh * HD_PAD:(hh + 1) * HD_PAD], scores, vt, n_store=HD_PAD, impl=4)
            h = self._linear(o, L["out"], r1=h)
            n2 = native.layernorm(self._new(B * Sp, C), h, *L["ln2"], eps=cfg["layer_norm_eps"])
            h = self._linear(self._linear(n2, L["fc1"]), L["fc2"], r1=h)
        hs = h.reshape(B, Sp, C)
        cls = hs[:, 0].contiguous()
        pooled = native.layernorm(self._new(B, C), cls, *self.post_ln, eps=cfg["layer_norm_eps"])
        emb = native.linear_small(self._new(B, self.proj.shape[0]), pooled, self.proj, None)
        return SimpleNamespace(image_embeds=emb, last_hidden_state=hs[:, :S])
