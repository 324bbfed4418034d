"""One tcgen05 GEMM shape/variant alone (for ncu): python tools/gemm_once.py --M 230400 --N 320 --K 2880 --impl 0"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=230400)
ap.add_argument("--N", type=int, default=320)
ap.add_argument("--K", type=int, default=2880)
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--no-res", action="store_true", help="no residual input (e.g. the qkv projection)")
ap.add_argument("--temporal", action="store_true", help="with --conv: the (3,1,1) temporal conv (K = 3*C) instead of 3x3")
ap.add_argument("--conv", type=int, nargs=5, default=None, metavar=("B", "F", "H", "W", "C"),
                help="3x3 conv over a channels-last activation instead of a plain matrix (K = 9*C)")
a = ap.parse_args()
if a.conv:
    B_, F_, H_, W_, C_ = a.conv
    a.M, a.K = B_ * F_ * H_ * W_, (3 if a.temporal else 9) * C_
bn = {3: 256, 5: 256, 4: 128, 7: 128, 6: 320}.get(a.impl, 160)
npad = (a.N + bn - 1) // bn * bn
w = torch.zeros(npad, a.K, device="cuda", dtype=torch.float16)
w[:a.N] = torch.randn(a.N, a.K, device="cuda").half() * a.K ** -0.5
b = torch.zeros(npad, device="cuda", dtype=torch.float16)
xs = [torch.randn(a.M, a.conv[4] if a.conv else a.K, device="cuda").half() for _ in range(2)]
r = None if a.no_res else torch.randn(a.M, a.N, device="cuda").half()
out = torch.empty(a.M, a.N, device="cuda", dtype=torch.float16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters):
    e0.record()
    if a.conv:
        native.gemm(out, xs[i % 2], w, bias=b, r1=r, conv_dims=tuple(a.conv), taps=native.TAPS_T3 if a.temporal else native.TAPS_3X3, n_store=a.N, impl=a.impl)
    else:
        native.gemm(out, xs[i % 2], w, bias=b, r1=r, n_store=a.N, impl=a.impl)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"gemm M={a.M} N={a.N} K={a.K} impl={a.impl}: {ms:.4f} ms {2.0 * a.M * a.N * a.K / ms / 1e9:.1f} TFLOP/s")
