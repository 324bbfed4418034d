"""Spatial FMHA alone at an SVD-XT shape (for ncu and quick timing).
  python tools/attn_once.py [--S 9216] [--heads 5] [--imgs 25] [--iters 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--S", type=int, default=9216)
ap.add_argument("--heads", type=int, default=5)
ap.add_argument("--imgs", type=int, default=25)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--impl", type=int, default=0)
a = ap.parse_args()
C = a.heads * 64
M = a.imgs * a.S
torch.manual_seed(0)
qkv = torch.randn(M, 3 * C, device="cuda", dtype=torch.float16)
out = torch.empty(M, C, device="cuda", dtype=torch.float16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(a.iters):
    e0.record()
    native.attn_spatial(out, qkv, n_img=a.imgs, S=a.S, heads=a.heads, q_off=0, k_off=C, v_off=2 * C, scale=0.125, impl=a.impl)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
fl = 4.0 * a.S * a.S * 64 * a.heads * a.imgs
print(f"attn_spatial impl={a.impl} S={a.S} heads={a.heads} imgs={a.imgs}: {best:.3f} ms  {fl / best / 1e9:.1f} TFLOP/s  finite={bool(torch.isfinite(out).all())}")
