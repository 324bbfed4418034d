"""Spatial FMHA variants x tile stagger at the SVD-XT shapes, in one process (A/B on the same box and clocks).
  python tools/attn_sweep.py [--impls 2,4,5,6] [--staggers 0,300,600,900,1200] [--out attn_sweep.json]
Each (impl, stagger) is timed as the best and the median of --iters launches after a warm-up, and its output is
compared with impl 2 (the round-1 kernel) so that a fast-but-wrong variant cannot slip through."""
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

ap = argparse.ArgumentParser()
ap.add_argument("--impls", default="2,4,5,6")
ap.add_argument("--staggers", default="0,300,600,900,1200")
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--key", default="fmha_stagger", help="tuning key the --staggers values are written to (fmha_stagger | fmha_handover)")
ap.add_argument("--out", default="attn_sweep.json")
ap.add_argument("--shapes", default="9216x5,2304x10", help="SxHEADS list (25 images each)")
a = ap.parse_args()
impls = [int(x) for x in a.impls.split(",")]
staggers = [int(x) for x in a.staggers.split(",")]
shapes = [(int(x.split("x")[0]), int(x.split("x")[1]), 25) for x in a.shapes.split(",")]
res = []
for S, heads, imgs in shapes:
    C = heads * 64
    M = imgs * S
    torch.manual_seed(0)
    qkv = torch.randn(M, 3 * C, device="cuda", dtype=torch.float16)
    ref = torch.empty(M, C, device="cuda", dtype=torch.float16)
    native.set_tuning("fmha_stagger", 0)
    native.attn_spatial(ref, qkv, n_img=imgs, S=S, heads=heads, q_off=0, k_off=C, v_off=2 * C, scale=0.125, impl=2)
    fl = 4.0 * S * S * 64 * heads * imgs
    for impl in impls:
        for st in staggers:
            native.set_tuning(a.key, st)
            out = torch.empty_like(ref)
            ts = []
            for i in range(a.iters + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                native.attn_spatial(out, qkv, n_img=imgs, S=S, heads=heads, q_off=0, k_off=C, v_off=2 * C, scale=0.125,
                                    impl=impl)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            err = (out.float() - ref.float()).abs().max().item()
            row = dict(S=S, heads=heads, imgs=imgs, impl=impl, key=a.key, stagger=st, best_ms=min(ts), median_ms=statistics.median(ts),
                       tflops_best=fl / min(ts) / 1e9, tflops_median=fl / statistics.median(ts) / 1e9,
                       max_abs_vs_impl2=err, finite=bool(torch.isfinite(out).all()))
            res.append(row)
            print(json.dumps(row), flush=True)
native.set_tuning("fmha_stagger", 900)
native.set_tuning("fmha_handover", 2)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", a.out), "w"), indent=1)
