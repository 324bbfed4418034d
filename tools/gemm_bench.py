"""A/B timing of the tcgen05 GEMM kernel variants on the SVD-XT layer shapes, in one process (same GPU,
same clocks): python tools/gemm_bench.py [geglu|linear] ...
Each line: shape, variant, ms, TFLOP/s.  Buffers rotate so inputs come from HBM, not L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402
from vdpp_b200.models.native_unet import interleave_geglu  # noqa: E402

DEV = "cuda"


def timed(run, n_sets, iters=12):
    for i in range(3):
        run(i % n_sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i % n_sets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def geglu_case(M, C, impls):
    inner = 4 * C
    w = torch.randn(2 * inner, C, device=DEV).half() * C ** -0.5
    b = torch.randn(2 * inner, device=DEV).half()
    n_sets = 3
    a = [torch.randn(M, C, device=DEV).half() for _ in range(n_sets)]
    out = [torch.empty(M, inner, device=DEV, dtype=torch.float16) for _ in range(n_sets)]
    for impl in impls:
        wi, bi, _ = interleave_geglu(w, b, half=128 if impl in (3, 5) else 80)
        ms = timed(lambda i: native.gemm(out[i], a[i], wi, bias=bi, geglu=True, n_store=inner, impl=impl), n_sets)
        print(f"geglu  M={M} N={2 * inner} K={C} impl={impl}: {ms:.4f} ms  {2.0 * M * 2 * inner * C / ms / 1e9:.1f} TFLOP/s", flush=True)


def linear_case(M, N, K, impls, residual=True):
    w = torch.randn(N, K, device=DEV).half() * K ** -0.5
    b = torch.randn(N, device=DEV).half()
    n_sets = 3
    a = [torch.randn(M, K, device=DEV).half() for _ in range(n_sets)]
    r = [torch.randn(M, N, device=DEV).half() for _ in range(n_sets)] if residual else None
    out = [torch.empty(M, N, device=DEV, dtype=torch.float16) for _ in range(n_sets)]
    for impl in impls:
        bn = {3: 256, 5: 256, 4: 128, 6: 320}.get(impl, 160)
        if N % bn:
            continue
        ms = timed(lambda i: native.gemm(out[i], a[i], w, bias=b, r1=r[i] if residual else None, n_store=N, impl=impl), n_sets)
        print(f"linear M={M} N={N} K={K} res={residual} impl={impl}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["geglu", "linear"]
    if "geglu" in what:
        for M, C in ((230400, 320), (57600, 640), (14400, 1280)):
            geglu_case(M, C, (3, 5))
    if "linear" in what:
        for M, N, K in ((230400, 320, 320), (230400, 320, 1280), (230400, 960, 320), (57600, 640, 2560), (57600, 640, 640),
                        (14400, 1280, 5120), (3600, 1280, 1280)):
            linear_case(M, N, K, (0, 2, 3, 4))
    if "qkv" in what:    # N = 960 / 1920 padded to a multiple of 256 for the CTA-pair kernel vs 128x160 tiles
        for M, N, K in ((230400, 960, 320), (57600, 1920, 640)):
            linear_case(M, N, K, (0,), residual=False)
            w = torch.zeros((N + 255) // 256 * 256, K, device=DEV, dtype=torch.float16)
            w[:N] = torch.randn(N, K, device=DEV).half() * K ** -0.5
            a = [torch.randn(M, K, device=DEV).half() for _ in range(3)]
            out = [torch.empty(M, N, device=DEV, dtype=torch.float16) for _ in range(3)]
            ms = timed(lambda i: native.gemm(out[i], a[i], w, n_store=N, impl=3), 3)
            print(f"linear M={M} N={N}->pad{w.shape[0]} K={K} impl=3: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    if "wide" in what:   # the 256x320 pair tile against the one-CTA 128x160 tile on the N = 320 / 640 layers
        for M, N, K in ((230400, 320, 2880), (230400, 320, 960), (230400, 320, 1280), (230400, 320, 640), (57600, 640, 5760),
                        (57600, 640, 2560), (57600, 640, 1920), (57600, 640, 640), (230400, 960, 320), (14400, 1920, 1280)):
            linear_case(M, N, K, (0, 6))
    if "sweep" in what:  # every tile shape on the N = 320 / 640 / 1280 layers (after an epilogue change the best pick may move)
        for M, N, K in ((230400, 320, 320), (230400, 320, 640), (230400, 320, 960), (230400, 320, 1280), (230400, 320, 2880),
                        (57600, 640, 640), (57600, 640, 1280), (57600, 640, 1920), (57600, 640, 2560), (57600, 640, 5760),
                        (14400, 1280, 1280), (14400, 1280, 3840), (3600, 1280, 1280), (3600, 1280, 3840), (3600, 1280, 11520)):
            linear_case(M, N, K, (0, 2, 3, 4, 6))
