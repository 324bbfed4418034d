"""Full-size SVD-XT UNet step on the native kernels: per-launch timing of the tcgen05 kernels (CUDA events)
and, with --library, the same step on torch's library kernels (the 'GPU library baseline' of BASELINE.md 4).
Writes gpurun_out/profile_forward.json.   python tools/profile_forward.py [--frames 25] [--library] [--once]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402
from vdpp_b200.models import StableVideoUNet  # noqa: E402
from vdpp_b200.models.native_unet import flops_per_forward  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--height", type=int, default=72)
    ap.add_argument("--width", type=int, default=128)
    ap.add_argument("--once", action="store_true", help="a single eager step and exit (for ncu)")
    ap.add_argument("--library", action="store_true", help="also time the torch oracle in fp16 on the GPU")
    ap.add_argument("--guidance-scale", type=float, default=None)
    ap.add_argument("--orchestrator", default="python", choices=["python", "c"],
                    help="per-launch CUDA events need the Python orchestration (one ctypes call per kernel); the C "
                         "orchestration (csrc/unet.cu, the product default) is timed per step beside it")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    F_, H, W = a.frames, a.height, a.width
    model = StableVideoUNet.from_pretrained("random-init:0", device=dev, orchestrator=a.orchestrator)
    torch.manual_seed(1)
    model.set_dummy_conditioning(1, F_, H, W, dev, guidance_scale=a.guidance_scale)
    x = torch.randn(1, 4, F_, H, W, device=dev).half() * model.init_noise_sigma
    out = model(x, 0)
    torch.cuda.synchronize()
    if a.once:
        print("finite", bool(torch.isfinite(out).all()), "absmax", out.float().abs().max().item())
        return
    res = {"frames": F_, "latent": [H, W], "finite": bool(torch.isfinite(out).all()), "orchestrator": a.orchestrator}
    if a.orchestrator == "python":      # the product path (one svdpp_unet_step call per step), eager and graph
        cm = StableVideoUNet.from_pretrained("random-init:0", device=dev, orchestrator="c")
        cm.set_conditioning(model._image_embeddings, model._image_latents, guidance_scale=a.guidance_scale, num_frames=F_)
        res["c_equals_python"] = bool(torch.equal(cm(x, 0), out))
        for mode in ("eager", "graph"):
            cm.use_cuda_graph = mode == "graph"
            for _ in range(3):
                cm(x, 1)
            torch.cuda.synchronize()
            tt = []
            for _ in range(5):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                cm(x, 1)
                c1.record()
                torch.cuda.synchronize()
                tt.append(c0.elapsed_time(c1))
            res[f"c_{mode}_step_ms"] = min(tt)
        del cm
        torch.cuda.empty_cache()
    # eager timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record()
        model(x, 1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res["eager_step_ms"] = min(ts)
    # graph timing
    model.use_cuda_graph = True
    for _ in range(3):
        model(x, 1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0.record()
        model(x, 1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res["graph_step_ms"] = min(ts)
    model.use_cuda_graph = False
    # per-launch profile
    native.PROFILE = []
    model(x, 1)
    torch.cuda.synchronize()
    prof, native.PROFILE = native.PROFILE, None
    rows = []
    agg = {}
    for kind, flops, shape, s, e in prof:
        ms = s.elapsed_time(e)
        rows.append(dict(kind=kind, shape=list(shape), ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1)))
        d = agg.setdefault(kind, [0.0, 0.0, 0])
        d[0] += ms
        d[1] += flops
        d[2] += 1
    res["by_kind"] = {k: dict(ms=round(v[0], 3), tflops=round(v[1] / v[0] / 1e9, 1), launches=v[2]) for k, v in agg.items()}
    fl = flops_per_forward(model.unet.cfg, 2 if a.guidance_scale else 1, F_, H, W)
    res["algorithmic_tflop"] = {k: round(v / 1e12, 3) for k, v in fl.items()}
    res["tc_ms_total"] = round(sum(v[0] for v in agg.values()), 3)
    res["other_ms"] = round(res["eager_step_ms"] - res["tc_ms_total"], 3)
    res["whole_step_tflops_graph"] = round(fl["total"] / res["graph_step_ms"] / 1e9, 1)
    # group identical shapes
    by_shape = {}
    for r in rows:
        key = (r["kind"], tuple(r["shape"]))
        d = by_shape.setdefault(key, [0.0, 0])
        d[0] += r["ms"]
        d[1] += 1
    top = sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:40]
    res["top_shapes"] = [dict(kind=k[0], shape=list(k[1]), total_ms=round(v[0], 3), n=v[1],
                              tflops=round((2.0 * k[1][0] * k[1][1] * k[1][2] if k[0] != "attn_spatial" else
                                            4.0 * k[1][1] ** 2 * 64 * k[1][2] * k[1][0]) * v[1] / v[0] / 1e9, 1))
                         for k, v in top]
    if a.library:
        from oracle.svd_step import OracleStep, Conditioning
        from oracle.unet_torch import UNetSpatioTemporalConditionModel
        del model
        torch.cuda.empty_cache()
        with torch.device(dev):
            lib = UNetSpatioTemporalConditionModel().half().eval()
        st = OracleStep(lib, 25)
        torch.manual_seed(1)
        cond = Conditioning(torch.randn(1, 1, 1024, device=dev).half(), torch.randn(1, 4, F_, H, W, device=dev).half(),
                            dtype=torch.float16, num_frames=F_, guidance_scale=a.guidance_scale)
        for _ in range(2):
            st(x, 1, cond)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0.record()
            st(x, 1, cond)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res["library_fp16_step_ms"] = min(ts)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "profile_forward.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "top_shapes"}, indent=1))
    for r in res["top_shapes"][:25]:
        print(r)


if __name__ == "__main__":
    main()
