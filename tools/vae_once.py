"""One eager NativeVAE decode of a chunk of frames at full size (latent 72x128 -> 576x1024), for `ncu` launch lists.
   python tools/vae_once.py [--frames 14] [--events]
--events prints per-kernel-family CUDA-event times instead (no profiler): GEMM / conv calls by shape, GroupNorm, the rest."""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from oracle.vae_torch import AutoencoderKLTemporalDecoder  # noqa: E402  (only its default-initialised state_dict)
from vdpp_b200 import native  # noqa: E402
from vdpp_b200.models.native_vae import NativeVAE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=14)
ap.add_argument("--events", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
lib = AutoencoderKLTemporalDecoder().half()
vae = NativeVAE(lib.state_dict(), device=dev)
del lib
lat = torch.randn(a.frames, 4, 72, 128, device=dev).half()
if not a.events:
    out = vae.decode(lat, num_frames=a.frames).sample
    torch.cuda.synchronize()
    print("decode ok", tuple(out.shape), bool(torch.isfinite(out).all()))
    sys.exit(0)

# per-call CUDA-event timing: wrap the native entry points the VAE uses
vae.decode(lat, num_frames=a.frames)
torch.cuda.synchronize()
records = []


def wrap(name, key_fn):
    orig = getattr(native, name)

    def f(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*args, **kw)
        e1.record()
        records.append((name, key_fn(args, kw), e0, e1))
        return r
    setattr(native, name, f)


def gemm_key(args, kw):
    out, x, w = args[0], args[1], args[2]
    cd = kw.get("conv_dims")
    taps = kw.get("taps")
    M = out.shape[0] if cd is None else cd[0] * cd[1] * cd[2] * cd[3]
    return f"M={M} N={kw.get('n_store', w.shape[0])} K={w.shape[1]} conv={0 if cd is None else len(taps) if taps is not None else 1}"


for nm in ("gemm",):
    wrap(nm, gemm_key)
wrap("groupnorm_silu", lambda args, kw: f"M={args[1].shape[0]} C={args[1].shape[1]} fps={kw.get('frames_per_stat', 1)}")
for nm in ("softmax_rows", "transpose", "time_conv_out", "pack_unet_input", "im2col", "upsample2x"):
    wrap(nm, lambda args, kw: "")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
vae.decode(lat, num_frames=a.frames)
e1.record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for name, key, s, e in records:
    agg[(name, key)][0] += s.elapsed_time(e)
    agg[(name, key)][1] += 1
total = e0.elapsed_time(e1)
rows = [dict(op=k[0], shape=k[1], ms=round(v[0], 3), n=v[1]) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])]
print(json.dumps(dict(frames=a.frames, decode_ms_with_events=total, covered_ms=sum(r["ms"] for r in rows), rows=rows[:40])))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(frames=a.frames, decode_ms_with_events=total, rows=rows), open(os.path.join(ROOT, "gpurun_out", "vae_once_events.json"), "w"), indent=1)
