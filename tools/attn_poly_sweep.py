"""Ping-pong FMHA (impl 7): share of the exponentials on the FMA pipe ("fmha_poly" of every 16) x hand-over point
("fmha_handover"), at the SVD-XT shapes, in one process (A/B on the same box and clocks).
  python tools/attn_poly_sweep.py [--polys 0,2,3,4,5,6,8] [--handovers 1,2,3] [--shapes 9216x5,2304x10] [--out attn_poly_sweep.json]
Each setting is timed as the best and the median of --iters launches after a warm-up; its output is compared with torch's
fp32 SDPA on a sample of images (max-abs and the error relative to the all-MUFU kernel's own error), so a fast-but-wrong
variant cannot slip through."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--polys", default="0,2,3,4,5,6,8")
ap.add_argument("--handovers", default="1,2,3")
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--out", default="attn_poly_sweep.json")
ap.add_argument("--shapes", default="9216x5,2304x10", help="SxHEADS list (25 images each)")
a = ap.parse_args()
polys = [int(x) for x in a.polys.split(",")]
handovers = [int(x) for x in a.handovers.split(",")]
shapes = [(int(x.split("x")[0]), int(x.split("x")[1]), 25) for x in a.shapes.split(",")]
old = {k: native.get_tuning(k) for k in ("fmha_poly", "fmha_handover")}
res = []
for S, heads, imgs in shapes:
    C = heads * 64
    M = imgs * S
    torch.manual_seed(0)
    qkv = torch.randn(M, 3 * C, device="cuda", dtype=torch.float16)
    # fp32 reference of image 0 (all heads)
    q, k, v = [t[:S].reshape(1, S, heads, 64).transpose(1, 2).float() for t in qkv.split(C, dim=1)]
    ref0 = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(S, C)
    fl = 4.0 * S * S * 64 * heads * imgs
    for poly in polys:
        for ho in handovers:
            native.set_tuning("fmha_poly", poly)
            native.set_tuning("fmha_handover", ho)
            out = torch.empty(M, C, device="cuda", dtype=torch.float16)
            ts = []
            for i in range(a.iters + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                native.attn_spatial(out, qkv, n_img=imgs, S=S, heads=heads, q_off=0, k_off=C, v_off=2 * C, scale=0.125, impl=7)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            err = (out[:S].float() - ref0).abs().max().item()
            row = dict(S=S, heads=heads, imgs=imgs, impl=7, fmha_poly=poly, fmha_handover=ho, best_ms=min(ts),
                       median_ms=statistics.median(ts), tflops_best=fl / min(ts) / 1e9, tflops_median=fl / statistics.median(ts) / 1e9,
                       max_abs_vs_fp32_sdpa=err, ref_absmax=ref0.abs().max().item(), finite=bool(torch.isfinite(out).all()))
            res.append(row)
            print(json.dumps(row), flush=True)
for k_, v_ in old.items():
    native.set_tuning(k_, v_)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", a.out), "w"), indent=1)
