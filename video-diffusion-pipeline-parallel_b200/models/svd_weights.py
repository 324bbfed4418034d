"""Parameter inventory of the SVD ``UNetSpatioTemporalConditionModel`` (diffusers ``state_dict`` keys
and shapes) and a seeded random-init generator for it.

The reference always loads pretrained weights from the hub (``src/models/svd_unet.py:129-136``); there is
no network here, and BASELINE.json asks for "random-init SVD weights", so benchmarks build the state
dict directly on the GPU with torch's default initialisers (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv
and linear weights and biases, ones/zeros for norms, 0.5 for the AlphaBlender mix factors).
``param_shapes`` must list exactly 1 524 623 082 parameters for the SVD config (tests check it against
the oracle module).  A real checkpoint in this key layout can be passed to NativeUNet unchanged.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from .native_unet import SVD_CONFIG


def param_shapes(config: Optional[dict] = None) -> "OrderedDict[str, Tuple[int, ...]]":
    cfg = dict(SVD_CONFIG)
    if config:
        cfg.update(config)
    boc = tuple(cfg["block_out_channels"])
    heads = tuple(cfg["num_attention_heads"])
    attn = tuple(cfg["down_attn"])
    L = cfg["layers_per_block"]
    xdim = cfg["cross_attention_dim"]
    temb = boc[0] * 4
    P: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def lin(name, i, o, bias=True):
        P[name + ".weight"] = (o, i)
        if bias:
            P[name + ".bias"] = (o,)

    def conv(name, i, o, k):
        P[name + ".weight"] = (o, i) + tuple(k)
        P[name + ".bias"] = (o,)

    def norm(name, c):
        P[name + ".weight"] = (c,)
        P[name + ".bias"] = (c,)

    def resblock(name, cin, cout):
        s, t = name + ".spatial_res_block", name + ".temporal_res_block"
        norm(s + ".norm1", cin); conv(s + ".conv1", cin, cout, (3, 3)); lin(s + ".time_emb_proj", temb, cout)
        norm(s + ".norm2", cout); conv(s + ".conv2", cout, cout, (3, 3))
        if cin != cout:
            conv(s + ".conv_shortcut", cin, cout, (1, 1))
        norm(t + ".norm1", cout); conv(t + ".conv1", cout, cout, (3, 1, 1)); lin(t + ".time_emb_proj", temb, cout)
        norm(t + ".norm2", cout); conv(t + ".conv2", cout, cout, (3, 1, 1))
        P[name + ".time_mixer.mix_factor"] = (1,)

    def attention(name, c, ctx):
        lin(name + ".to_q", c, c, False); lin(name + ".to_k", ctx, c, False); lin(name + ".to_v", ctx, c, False)
        lin(name + ".to_out.0", c, c)

    def ff(name, c):
        lin(name + ".net.0.proj", c, 8 * c); lin(name + ".net.2", 4 * c, c)

    def transformer(name, c):
        norm(name + ".norm", c); lin(name + ".proj_in", c, c)
        s = name + ".transformer_blocks.0"
        norm(s + ".norm1", c); attention(s + ".attn1", c, c); norm(s + ".norm2", c); attention(s + ".attn2", c, xdim)
        norm(s + ".norm3", c); ff(s + ".ff", c)
        t = name + ".temporal_transformer_blocks.0"
        norm(t + ".norm_in", c); ff(t + ".ff_in", c)
        norm(t + ".norm1", c); attention(t + ".attn1", c, c); norm(t + ".norm2", c); attention(t + ".attn2", c, xdim)
        norm(t + ".norm3", c); ff(t + ".ff", c)
        lin(name + ".time_pos_embed.linear_1", c, 4 * c); lin(name + ".time_pos_embed.linear_2", 4 * c, c)
        P[name + ".time_mixer.mix_factor"] = (1,)
        lin(name + ".proj_out", c, c)

    conv("conv_in", cfg["in_channels"], boc[0], (3, 3))
    lin("time_embedding.linear_1", boc[0], temb); lin("time_embedding.linear_2", temb, temb)
    lin("add_embedding.linear_1", cfg["projection_class_embeddings_input_dim"], temb)
    lin("add_embedding.linear_2", temb, temb)
    c = boc[0]
    for i, co in enumerate(boc):
        for j in range(L):
            resblock(f"down_blocks.{i}.resnets.{j}", c, co)
            c = co
        if attn[i]:
            for j in range(L):
                transformer(f"down_blocks.{i}.attentions.{j}", co)
        if i != len(boc) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", co, co, (3, 3))
    resblock("mid_block.resnets.0", c, c); resblock("mid_block.resnets.1", c, c)
    transformer("mid_block.attentions.0", c)
    rev, rattn = boc[::-1], attn[::-1]
    out_c = rev[0]
    for i, co in enumerate(rev):
        prev_c, out_c = out_c, co
        in_c = rev[min(i + 1, len(boc) - 1)]
        for j in range(L + 1):
            skip_c = in_c if j == L else out_c
            res_in = prev_c if j == 0 else out_c
            resblock(f"up_blocks.{i}.resnets.{j}", res_in + skip_c, out_c)
        if rattn[i]:
            for j in range(L + 1):
                transformer(f"up_blocks.{i}.attentions.{j}", out_c)
        if i != len(boc) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, (3, 3))
    norm("conv_norm_out", boc[0])
    conv("conv_out", boc[0], cfg["out_channels"], (3, 3))
    return P


def param_count(config: Optional[dict] = None) -> int:
    return sum(math.prod(s) for s in param_shapes(config).values())


def random_state_dict(config: Optional[dict] = None, seed: int = 0, device="cuda",
                      dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Default-initialiser weights drawn from one seeded generator on ``device`` (deterministic per
    (seed, device type); not the same stream of numbers as constructing the torch module)."""
    shapes = param_shapes(config)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    fan_in: Dict[str, int] = {}
    for name, shp in shapes.items():
        if name.endswith(".weight") and len(shp) >= 2:
            fan_in[name[:-len(".weight")]] = math.prod(shp[1:])
    for name, shp in shapes.items():
        base = name.rsplit(".", 1)[0]
        if name.endswith("mix_factor"):
            t = torch.full(shp, 0.5, device=device, dtype=torch.float32)
        elif base in fan_in:
            bound = 1.0 / math.sqrt(fan_in[base])
            t = (torch.rand(shp, generator=gen, device=device, dtype=torch.float32) * 2 - 1) * bound
        elif name.endswith(".weight"):
            t = torch.ones(shp, device=device, dtype=torch.float32)
        else:
            t = torch.zeros(shp, device=device, dtype=torch.float32)
        sd[name] = t.to(dtype)
    return sd
