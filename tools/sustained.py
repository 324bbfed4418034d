"""Sustained (power-capped) throughput and power draw per kernel class of the SVD step.

Short A/B loops run at boost clocks (1965 MHz, < 500 W) and overstate what a kernel delivers inside the seconds-long
denoising run, where the board sits at its 1000 W cap near 1.55-1.6 GHz.  This tool loops ONE kernel for `--secs`
seconds, times the last half with CUDA events and samples NVML power / SM clock meanwhile: throughput at the cap, watts,
and therefore joules per unit of work - the quantity that bounds the step once the cap is reached.
   python tools/sustained.py [--secs 2.0] [name-substring ...]    ->  gpurun_out/sustained.jsonl
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402
from vdpp_b200.models.native_unet import interleave_geglu  # noqa: E402

DEV = "cuda"


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.on = False
        self.stop = False
        self.p, self.c = [], []

    def run(self):
        while not self.stop:
            if self.on:
                try:
                    self.p.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                    self.c.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.02)


def h(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).half()


def cases(frames):
    M0, M1, M2 = frames * 9216, frames * 2304, frames * 576
    out = {}

    def gemm_case(name, M, N, K, impl, residual=False, geglu=False):
        def make():
            x = h(M, K, scale=0.5)
            w = h(N, K, scale=K ** -0.5)
            b = h(N, scale=0.1)
            n_out = N
            if geglu:
                w, b, _ = interleave_geglu(w, b, half=128)
                n_out = N // 2
            r1 = h(M, n_out) if residual else None
            o = torch.empty(M, n_out, device=DEV, dtype=torch.float16)
            return lambda: native.gemm(o, x, w, bias=b, r1=r1, geglu=geglu, n_store=n_out, impl=impl)
        out[name] = (make, 2.0 * M * N * K, "TFLOP")

    gemm_case("geglu_L0_K320", M0, 2560, 320, 3, geglu=True)
    gemm_case("ff2_L0_K1280", M0, 320, 1280, 6, residual=True)
    gemm_case("qkv_L0_K320", M0, 1024, 320, 3)
    gemm_case("proj_L0_K320_res", M0, 320, 320, 0, residual=True)
    gemm_case("linear_L2_K5120", M2 * 4, 1280, 5120, 3, residual=True)
    gemm_case("geglu_L2_K1280", M2 * 4, 10240, 1280, 3, geglu=True)
    gemm_case("gemm_L0_K2880_pair320", M0, 320, 2880, 6)
    gemm_case("gemm_L1_K5760_pair320", M1, 640, 5760, 6)

    def ff_case(name, M, Cc, fused):
        def make():
            inner = 4 * Cc
            x = h(M, Cc, scale=0.5)
            w1 = h(2 * inner, Cc, scale=Cc ** -0.5)
            b1 = h(2 * inner, scale=0.1)
            w2 = h(Cc, inner, scale=inner ** -0.5)
            b2 = h(Cc, scale=0.1)
            r1 = h(M, Cc)
            o = torch.empty(M, Cc, device=DEV, dtype=torch.float16)
            if fused:
                w1i, b1i, _ = interleave_geglu(w1, b1, half=64)
                return lambda: native.ff_geglu(o, x, w1i, b1i, w2, b2, r1=r1)
            w1i, b1i, _ = interleave_geglu(w1, b1, half=128)
            mid = torch.empty(M, inner, device=DEV, dtype=torch.float16)

            def run():
                native.gemm(mid, x, w1i, bias=b1i, geglu=True, n_store=inner, impl=3)
                native.gemm(o, mid, w2, bias=b2, r1=r1, n_store=Cc, impl=6)
            return run
        out[name] = (make, 2.0 * M * Cc * 12 * Cc, "TFLOP")

    ff_case("ff_L0_two_kernels", M0, 320, False)
    ff_case("ff_L0_fused", M0, 320, True)

    def cublas_case(name, M, N, K):   # the library's kernel at the same shape: what does a joule buy there?
        def make():
            x = h(M, K, scale=0.5)
            w = h(N, K, scale=K ** -0.5)
            o = torch.empty(M, N, device=DEV, dtype=torch.float16)
            return lambda: torch.matmul(x, w.t(), out=o)
        out[name] = (make, 2.0 * M * N * K, "TFLOP")

    cublas_case("cublas_L0_K2880", M0, 320, 2880)
    cublas_case("cublas_L1_K5760", M1, 640, 5760)
    cublas_case("cublas_L2_K5120", M2 * 4, 1280, 5120)
    cublas_case("cublas_L0_N2560_K320", M0, 2560, 320)
    cublas_case("cublas_8192cube", 8192, 8192, 8192)

    def attn_case(name, S, heads, impl):
        def make():
            C = heads * 64
            qkv = h(frames * S, 3 * C)
            o = torch.empty(frames * S, C, device=DEV, dtype=torch.float16)
            return lambda: native.attn_spatial(o, qkv, n_img=frames, S=S, heads=heads, q_off=0, k_off=C, v_off=2 * C,
                                               scale=0.125, impl=impl)
        out[name] = (make, 4.0 * S * S * 64 * heads * frames, "TFLOP")

    attn_case("attn_S9216_impl4", 9216, 5, 4)
    attn_case("attn_S9216_impl7", 9216, 5, 7)

    def gn_case(name, n_img, HW, C):
        def make():
            x = h(n_img * HW, C)
            g, b = h(C), h(C)
            o = torch.empty_like(x)
            ws = torch.empty(native.groupnorm_workspace_bytes(n_img, HW), device=DEV, dtype=torch.uint8)
            return lambda: native.groupnorm_silu(o, x, g, b, n_img=n_img, HW=HW, eps=1e-5, workspace=ws)
        out[name] = (make, 2.0 * n_img * HW * C * 2, "GB")

    gn_case("groupnorm_L0", frames, 9216, 320)

    def ln_case(name, M, C):
        def make():
            x = h(M, C)
            g, b = h(C), h(C)
            o = torch.empty_like(x)
            return lambda: native.layernorm(o, x, g, b)
        out[name] = (make, 2.0 * M * C * 2, "GB")

    ln_case("layernorm_L0", M0, 320)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--secs", type=float, default=2.0)
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("filters", nargs="*")
    a = ap.parse_args()
    smp = Sampler()
    smp.start()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "sustained.jsonl"), "a")
    for name, (make, work, unit) in cases(a.frames).items():
        if a.filters and not any(f in name for f in a.filters):
            continue
        run = make()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        burst_ms = e0.elapsed_time(e1) / 10
        n_half = max(10, int(a.secs * 500.0 / burst_ms))
        for _ in range(n_half):  # first half: heat up (untimed)
            run()
        e0.record()
        torch.cuda.synchronize()
        smp.p, smp.c = [], []
        smp.on = True
        e0.record()
        for _ in range(n_half):
            run()
        e1.record()
        torch.cuda.synchronize()
        smp.on = False
        sus_ms = e0.elapsed_time(e1) / n_half
        watts = statistics.median(smp.p) if smp.p else None
        row = dict(name=name, burst_ms=round(burst_ms, 4), sustained_ms=round(sus_ms, 4), slowdown=round(sus_ms / burst_ms, 3),
                   burst_rate=round(work / burst_ms / 1e9, 1), sustained_rate=round(work / sus_ms / 1e9, 1), unit=unit + "/s",
                   watts=watts, sm_mhz=statistics.median(smp.c) if smp.c else None,
                   joule_per_unit=round(watts * sus_ms * 1e-3 / (work / 1e12 if unit == "TFLOP" else work / 1e9), 4) if watts else None,
                   samples=len(smp.p))
        print(json.dumps(row), flush=True)
        log.write(json.dumps(row) + "\n")
        del run
        torch.cuda.empty_cache()
        time.sleep(1.0)   # cool down between cases
    smp.stop = True


if __name__ == "__main__":
    main()
