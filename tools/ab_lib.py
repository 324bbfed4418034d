"""A/B of two BUILDS of libsvdpp.so inside one process (box-to-box and minute-to-minute clock drift under the power
cap is larger than the effects being measured, so both builds alternate on the same GPU):
    python tools/ab_lib.py --a gpurun_out/libsvdpp_a.so [--b <csrc/libsvdpp.so>] [--rounds 5] [--no-step]
Per GEMM shape of the network: median microseconds of each build over `rounds` alternating measurements, and the
full-size UNet step (eager) the same way.  Writes gpurun_out/ab_lib.json.
"""
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
from vdpp_b200.models.native_unet import interleave_geglu  # noqa: E402

SHAPES = [  # (M, N, K, impl, residual, geglu)
    (230400, 320, 320, 0, True, False), (230400, 320, 320, 0, False, False), (230400, 1024, 320, 3, False, False),
    (230400, 320, 1280, 6, True, False), (57600, 640, 640, 0, True, False), (57600, 2048, 640, 3, False, False),
    (57600, 640, 2560, 6, True, False), (14400, 1280, 1280, 3, True, False), (14400, 1280, 5120, 3, True, False),
    (230400, 320, 2880, 6, False, False), (57600, 640, 5760, 6, False, False), (230400, 320, 5760, 6, True, False),
    (230400, 2560, 320, 3, False, True), (57600, 5120, 640, 3, False, True),
    (14400, 10240, 1280, 3, False, True),
]
SHAPES_PAIR160 = [  # the N = 320 / 640 short-K layers on the 256x160 pair tile (impl 2) instead of the one-CTA 128x160 tile
    (230400, 320, 320, 2, True, False), (230400, 320, 320, 2, False, False), (57600, 640, 640, 2, True, False),
    (230400, 320, 640, 2, True, False), (230400, 320, 320, 0, True, False), (57600, 640, 640, 0, True, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--a", required=True)
    ap.add_argument("--b", default=str(native.library_path()))
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-step", action="store_true")
    ap.add_argument("--shapes", default="net", choices=["net", "pair160"])
    ap.add_argument("--tune-a", default="", help="key=value[,key=value] tuning switches set on build a")
    ap.add_argument("--tune-b", default="", help="... on build b (a and b may then be copies of the same build)")
    a = ap.parse_args()
    tune = {"a": [kv.split("=") for kv in a.tune_a.split(",") if kv], "b": [kv.split("=") for kv in a.tune_b.split(",") if kv]}

    def use(k):
        native.use_library(libs[k])
        for key, val in tune[k]:
            native.set_tuning(key, int(val))
    dev = torch.device("cuda", 0)
    libs = {"a": a.a, "b": a.b}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {"a": a.a, "b": a.b, "tune": tune, "gemm": []}
    for M, N, K, impl, has_res, geglu in (SHAPES if a.shapes == "net" else SHAPES_PAIR160):
        x = (torch.randn(M, K, device=dev) * 0.5).half()
        if geglu:
            w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
            bias = (torch.randn(N, device=dev) * 0.1).half()
            w, bias, _ = interleave_geglu(w, bias, half=128)
            n_out = N // 2
        else:
            w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
            bias = torch.randn(N, device=dev).half()
            n_out = N
        r1 = torch.randn(M, n_out, device=dev).half() if has_res else None
        out = {k: torch.empty(M, n_out, device=dev, dtype=torch.float16) for k in libs}
        ts = {k: [] for k in libs}
        for rnd in range(a.rounds + 1):
            for k, path in libs.items():
                use(k)
                native.gemm(out[k], x, w, bias=bias, r1=r1, geglu=geglu, n_store=n_out, impl=impl)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.iters):
                    native.gemm(out[k], x, w, bias=bias, r1=r1, geglu=geglu, n_store=n_out, impl=impl)
                e1.record()
                torch.cuda.synchronize()
                if rnd:
                    ts[k].append(e0.elapsed_time(e1) / a.iters * 1e3)
        ma, mb = statistics.median(ts["a"]), statistics.median(ts["b"])
        row = dict(M=M, N=N, K=K, impl=impl, residual=has_res, geglu=geglu, a_us=round(ma, 1), b_us=round(mb, 1),
                   b_over_a_speedup=round(ma / mb, 3), b_tflops=round(2.0 * M * N * K / mb / 1e6, 1),
                   identical=bool(torch.equal(out["a"], out["b"])))
        res["gemm"].append(row)
        print(row, flush=True)
        del x, w, r1, out
    if not a.no_step:
        native.use_library(a.b)
        model = StableVideoUNet.from_pretrained("random-init:0", device=dev)
        torch.manual_seed(1)
        model.set_dummy_conditioning(1, 25, 72, 128, dev)
        x = torch.randn(1, 4, 25, 72, 128, device=dev).half() * model.init_noise_sigma
        ts = {k: [] for k in libs}
        outs = {}
        for rnd in range(a.rounds + 1):
            for k, path in libs.items():
                use(k)
                torch.cuda.synchronize()
                e0.record()
                y = x
                for s in range(3):
                    y = model(y, s)
                e1.record()
                torch.cuda.synchronize()
                outs[k] = y.clone()
                if rnd:
                    ts[k].append(e0.elapsed_time(e1) / 3)
        res["step_ms"] = {k: [round(t, 2) for t in v] for k, v in ts.items()}
        res["step_ms_median"] = {k: round(statistics.median(v), 2) for k, v in ts.items()}
        res["step_identical"] = bool(torch.equal(outs["a"], outs["b"]))
        print("step", res["step_ms"], res["step_ms_median"], "identical", res["step_identical"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_lib.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
