"""Parameter inventories, seeded random-init generators and local-checkpoint readers for the front / back end of the
image -> video run: the SVD ``AutoencoderKLTemporalDecoder`` (``vae/``) and the CLIP ViT-H/14 image encoder
(``image_encoder/``), next to ``svd_weights.py`` which does the same for the UNet.

The reference loads all three from the hub (``scripts/generate_video_demo.py:249-275``: ``from_pretrained(model_id,
subfolder=...)``).  There is no network here, so ``model_id`` is a LOCAL snapshot directory in the hub layout

    <dir>/unet/diffusion_pytorch_model[.fp16].safetensors
    <dir>/vae/diffusion_pytorch_model[.fp16].safetensors           (+ config.json)
    <dir>/image_encoder/model[.fp16].safetensors                   (+ config.json)
    <dir>/feature_extractor/preprocessor_config.json

or ``random-init[:seed]`` for default-initialised weights of the real architectures (benchmarks, tests).  The key
names are the diffusers / transformers ``state_dict`` keys, so a real checkpoint and a generated one go through the
same packing code of ``NativeVAE`` / ``NativeCLIPVision``.  tests/test_frontend_host.py checks both inventories against
the torch restatement (VAE: 97 742 847 parameters, the published size) and the real transformers class (CLIP).
"""
from __future__ import annotations

import json
import math
import os
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

Shapes = "OrderedDict[str, Tuple[int, ...]]"


# ---------------------------------------------------------------------------------------------- inventories
def vae_param_shapes(config: Optional[dict] = None) -> Shapes:
    from .native_vae import VAE_CONFIG
    cfg = dict(VAE_CONFIG)
    if config:
        cfg.update({k: v for k, v in config.items() if k in cfg})
    boc = tuple(cfg["block_out_channels"])
    L, lat = cfg["layers_per_block"], cfg["latent_channels"]
    P: Shapes = OrderedDict()

    def conv(name, i, o, k):
        P[name + ".weight"] = (o, i) + tuple(k)
        P[name + ".bias"] = (o,)

    def norm(name, c):
        P[name + ".weight"] = (c,)
        P[name + ".bias"] = (c,)

    def lin(name, i, o):
        P[name + ".weight"] = (o, i)
        P[name + ".bias"] = (o,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, (3, 3))
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, (3, 3))
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, (1, 1))

    def st_resblock(name, cin, cout):
        resnet(name + ".spatial_res_block", cin, cout)
        t = name + ".temporal_res_block"
        norm(t + ".norm1", cout); conv(t + ".conv1", cout, cout, (3, 1, 1))
        norm(t + ".norm2", cout); conv(t + ".conv2", cout, cout, (3, 1, 1))
        P[name + ".time_mixer.mix_factor"] = (1,)

    def attention(name, c):
        norm(name + ".group_norm", c)
        for p in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{name}.{p}", c, c)

    # encoder: plain AutoencoderKL encoder
    conv("encoder.conv_in", cfg["in_channels"], boc[0], (3, 3))
    c = boc[0]
    for i, co in enumerate(boc):
        for j in range(L):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, (3, 3))
    resnet("encoder.mid_block.resnets.0", c, c)
    attention("encoder.mid_block.attentions.0", c)
    resnet("encoder.mid_block.resnets.1", c, c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", c, 2 * lat, (3, 3))
    # temporal decoder
    rev = boc[::-1]
    conv("decoder.conv_in", lat, rev[0], (3, 3))
    for j in range(L):
        st_resblock(f"decoder.mid_block.resnets.{j}", rev[0], rev[0])
    attention("decoder.mid_block.attentions.0", rev[0])
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(L + 1):
            st_resblock(f"decoder.up_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i != len(boc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, (3, 3))
    norm("decoder.conv_norm_out", c)
    conv("decoder.conv_out", c, cfg["out_channels"], (3, 3))
    conv("decoder.time_conv_out", cfg["out_channels"], cfg["out_channels"], (3, 1, 1))
    conv("quant_conv", 2 * lat, 2 * lat, (1, 1))
    return P


def clip_param_shapes(config: Optional[dict] = None) -> Shapes:
    from .native_clip import CLIP_VIT_H
    cfg = dict(CLIP_VIT_H)
    if config:
        cfg.update({k: v for k, v in config.items() if k in cfg})
    C, I, p = cfg["hidden_size"], cfg["intermediate_size"], cfg["patch_size"]
    S = (cfg["image_size"] // p) ** 2 + 1
    P: Shapes = OrderedDict()

    def lin(name, i, o, bias=True):
        P[name + ".weight"] = (o, i)
        if bias:
            P[name + ".bias"] = (o,)

    def norm(name):
        P[name + ".weight"] = (C,)
        P[name + ".bias"] = (C,)

    v = "vision_model"
    P[f"{v}.embeddings.class_embedding"] = (C,)
    P[f"{v}.embeddings.patch_embedding.weight"] = (C, 3, p, p)
    P[f"{v}.embeddings.position_embedding.weight"] = (S, C)
    norm(f"{v}.pre_layrnorm")          # the spelling of the transformers key
    for i in range(cfg["num_hidden_layers"]):
        layer = f"{v}.encoder.layers.{i}"
        for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(f"{layer}.self_attn.{proj}", C, C)
        norm(f"{layer}.layer_norm1")
        lin(f"{layer}.mlp.fc1", C, I)
        lin(f"{layer}.mlp.fc2", I, C)
        norm(f"{layer}.layer_norm2")
    norm(f"{v}.post_layernorm")
    lin("visual_projection", C, cfg["projection_dim"], bias=False)
    return P


def param_count(shapes: Shapes) -> int:
    return sum(math.prod(s) for s in shapes.values())


# ---------------------------------------------------------------------------------------------- random init
def random_state_dict(shapes: Shapes, seed: int = 0, device="cuda", dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Default initialisers from one seeded generator on ``device``: U(+-1/sqrt(fan_in)) for conv / linear weights and
    their biases, ones / zeros for norms, 0.5 for the AlphaBlender mix factors, N(0, 0.02) for embedding tables."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    fan_in = {n[:-len(".weight")]: math.prod(s[1:]) for n, s in shapes.items() if n.endswith(".weight") and len(s) >= 2}
    sd: Dict[str, torch.Tensor] = {}
    for name, shp in shapes.items():
        base = name.rsplit(".", 1)[0]
        if name.endswith("mix_factor"):
            t = torch.full(shp, 0.5, device=device, dtype=torch.float32)
        elif name.endswith("class_embedding") or name.endswith("position_embedding.weight"):
            t = torch.randn(shp, generator=gen, device=device, dtype=torch.float32) * 0.02
        elif base in fan_in:
            bound = 1.0 / math.sqrt(fan_in[base])
            t = (torch.rand(shp, generator=gen, device=device, dtype=torch.float32) * 2 - 1) * bound
        elif name.endswith(".weight"):
            t = torch.ones(shp, device=device, dtype=torch.float32)
        else:
            t = torch.zeros(shp, device=device, dtype=torch.float32)
        sd[name] = t.to(dtype)
    return sd


# ---------------------------------------------------------------------------------------------- local checkpoints
def is_random_init(model_id: str) -> bool:
    return model_id.startswith("random-init")


def random_init_seed(model_id: str) -> int:
    return int(model_id.split(":", 1)[1]) if ":" in model_id else 0


def component_dir(model_id: str, subfolder: str) -> str:
    """``<model_id>/<subfolder>`` of a local snapshot (or ``model_id`` itself when it already is that folder)."""
    if not os.path.isdir(model_id):
        raise FileNotFoundError(f"'{model_id}' is not a local checkpoint directory and there is no network access; "
                                "use a local path or 'random-init[:seed]'")
    sub = os.path.join(model_id, subfolder)
    return sub if os.path.isdir(sub) else model_id


def load_component(model_id: str, subfolder: str, device="cuda") -> Tuple[Dict[str, torch.Tensor], Optional[dict]]:
    """(state dict, config.json or None) of one component; prefers the ``fp16`` variant, merges sharded files."""
    from safetensors.torch import load_file
    d = component_dir(model_id, subfolder)
    files = sorted(f for f in os.listdir(d) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors file under {d}")
    pick = [f for f in files if "fp16" in f] or files
    sharded = [f for f in pick if "-of-" in f]
    sd: Dict[str, torch.Tensor] = {}
    for f in (sharded or pick[:1]):
        sd.update(load_file(os.path.join(d, f), device=str(device)))
    cfg_path = os.path.join(d, "config.json")
    config = None
    if os.path.isfile(cfg_path):
        with open(cfg_path) as fh:
            config = json.load(fh)
    return sd, config
