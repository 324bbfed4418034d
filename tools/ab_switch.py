"""Generic in-process A/B of tuning switches on the full-size SVD-XT denoising step (25 frames, 72x128 latent).
   python tools/ab_switch.py zigzag=0,1 fmha_stagger=0,900 [--rounds 3] [--steps 4] [--attn-impl-long 4] [--graph]
Every combination of the listed values is timed over `--rounds` interleaved rounds of `--steps` steps (so clock drift
under the power cap hits all settings alike); outputs of all settings are compared bit for bit with the first one.
Writes gpurun_out/ab_switch.json and prints one JSON line per setting."""
import argparse
import itertools
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
ap.add_argument("switches", nargs="+", help="key=v0,v1,...")
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--frames", type=int, default=25)
ap.add_argument("--attn-impl-long", type=int, default=None)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--out", default="ab_switch.json")
a = ap.parse_args()
keys = [s.split("=")[0] for s in a.switches]
vals = [[int(v) for v in s.split("=")[1].split(",")] for s in a.switches]
dev = torch.device("cuda", 0)
F_, H, W = a.frames, 72, 128
model = StableVideoUNet(unet=NativeUNet(random_state_dict(None, seed=0, device=dev), device=dev,
                                        attn_impl_long=a.attn_impl_long),
                        timesteps=StableVideoUNet._default_timestep_schedule(25)).to(dev)
torch.manual_seed(1)
model.set_dummy_conditioning(1, F_, H, W, dev)
x = torch.randn(1, 4, F_, H, W, device=dev).half() * model.init_noise_sigma
defaults = {k: native.get_tuning(k) for k in keys}
combos = list(itertools.product(*vals))


def apply(c):
    for k, v in zip(keys, c):
        native.set_tuning(k, v)


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    y = x
    for s in range(n):
        y = model(y, s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, y


ref, same, times = None, {}, {c: [] for c in combos}
model.use_cuda_graph = a.graph
for c in combos:
    apply(c)
    if a.graph:
        model._graphs.clear()
        run(a.steps)
    _, y = run(a.steps)
    ref = y.clone() if ref is None else ref
    same[c] = bool(torch.equal(ref, y))
for r in range(a.rounds):
    for c in combos:
        apply(c)
        if a.graph:
            model._graphs.clear()
            run(a.steps)
        ms, _ = run(a.steps)
        times[c].append(ms)
for k, v in defaults.items():
    native.set_tuning(k, v)
res = []
for c in combos:
    row = dict(zip(keys, c))
    row.update(step_ms_median=round(statistics.median(times[c]), 3), step_ms_min=round(min(times[c]), 3),
               step_ms_all=[round(t, 3) for t in times[c]], identical_to_first=same[c], graph=a.graph)
    res.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", a.out), "w"), indent=1)
