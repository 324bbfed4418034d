"""In-situ A/B of the spatial FMHA variants: the full-size SVD-XT denoising step (25 frames, 72x128) with the UNet's
long-sequence attention on impl 2 (round 1) vs impl 4/5/6 (quarter-pipelined softmax) and tile stagger values,
alternating inside one process so that clock / power drift cancels.  Writes gpurun_out/ab_attn_insitu.json.
   python tools/ab_attn_insitu.py [--impls 2,4] [--staggers 0,900] [--rounds 3] [--steps 4]"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402
from vdpp_b200.models import StableVideoUNet  # noqa: E402
from vdpp_b200.models.native_unet import NativeUNet  # noqa: E402
from vdpp_b200.models.svd_weights import random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impls", default="2,4")
ap.add_argument("--staggers", default="0,900")
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--frames", type=int, default=25)
a = ap.parse_args()
dev = torch.device("cuda", 0)
F_, H, W = a.frames, 72, 128
sd = random_state_dict(None, seed=0, device=dev)
models = {}
torch.manual_seed(1)
emb = torch.randn(1, 1, 1024, device=dev).half()
img = torch.randn(1, 4, F_, H, W, device=dev).half()
for impl in [int(x) for x in a.impls.split(",")]:
    m = StableVideoUNet(unet=NativeUNet(sd, device=dev, attn_impl_long=impl),
                        timesteps=StableVideoUNet._default_timestep_schedule(25)).to(dev)
    m.set_conditioning(emb, img, num_frames=F_)
    models[impl] = m
del sd
x = torch.randn(1, 4, F_, H, W, device=dev).half() * 700.0
configs = [(i, s) for i in models for s in [int(v) for v in a.staggers.split(",")]]
ref = None
times = {c: [] for c in configs}
outs = {}
for r in range(a.rounds + 1):
    for c in configs:
        impl, st = c
        native.set_tuning("fmha_stagger", st)
        m = models[impl]
        y = m(x, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_ in range(a.steps):
            y = m(x, s_ % 25)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            times[c].append(e0.elapsed_time(e1) / a.steps)
        outs[c] = m(x, 0)
native.set_tuning("fmha_stagger", 0)
base = outs[configs[0]].float()
res = []
for c in configs:
    row = dict(attn_impl_long=c[0], stagger=c[1], step_ms_median=statistics.median(times[c]), step_ms_min=min(times[c]),
               max_abs_vs_first=(outs[c].float() - base).abs().max().item(), latent_absmax=base.abs().max().item())
    res.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_attn_insitu.json"), "w"), indent=1)
