"""What does the last, partial wave of the 256x320 pair tiles cost, and what does the split-K tail recover?
Times impl 6 at the network's M (partial last wave) without / with the split-K tail, and at M truncated to whole waves.
    python tools/tail_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402

dev = "cuda"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(M, N, K, res, split):
    native.set_tuning("splitk", split)
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
    b = torch.randn(N, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    best = 1e9
    for _ in range(4):
        for _ in range(2):
            native.gemm(out, a, w, bias=b, r1=r, n_store=N, impl=6)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            native.gemm(out, a, w, bias=b, r1=r, n_store=N, impl=6)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10 * 1e3)
    return best


pairs_sm = native.device_info()[2] // 2
for M, N, K, res in ((230400, 320, 2880, False), (230400, 320, 1280, True), (230400, 320, 960, False),
                     (57600, 640, 5760, False), (57600, 640, 2560, True), (57600, 640, 1920, False)):
    nt = N // 320
    pairs = (M // 256) * nt
    waves = pairs // pairs_sm
    m_trunc = waves * pairs_sm // nt * 256
    t_no, t_sk, t_tr = timed(M, N, K, res, 0), timed(M, N, K, res, 1), timed(m_trunc, N, K, res, 0)
    print(f"M={M} N={N} K={K} res={res}: {pairs} pair tiles = {waves} waves + {pairs - waves * pairs_sm}; "
          f"no split {t_no:.1f} us, split-K tail {t_sk:.1f} us, whole waves only (M={m_trunc}) {t_tr:.1f} us "
          f"-> tail costs {t_no - t_tr:.1f} us, split-K tail costs {t_sk - t_tr:.1f} us", flush=True)
native.set_tuning("splitk", 1)
