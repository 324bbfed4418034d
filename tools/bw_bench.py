"""Achieved HBM bandwidth of the bandwidth-bound kernels of the SVD step at the shapes of the SVD-XT
workload (25 frames, latent 72x128), against MEASURED_PEAKS.json's copy bandwidth.

For every kernel: ALGORITHMIC bytes (the tensors it must read and write once) / average launch time,
timed with CUDA events over a rotation of buffer sets larger than the 126 MB L2 (so no launch finds its
input in cache).  One JSON line per kernel on stdout and in gpurun_out/bw_bench.jsonl.
Usage (GPU box):  python tools/bw_bench.py [--frames 25] [name-substring ...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import native  # noqa: E402

DEV = "cuda"
L2_BYTES = 126e6


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        return 6550.0, "fallback (B200_PROFILING.md)"


def timed(make_set, run, set_bytes, iters=20):
    """make_set() -> one set of tensors; run(set) launches the kernel once.  Rotates >= 3 sets / > 4x L2."""
    n_sets = max(3, int(4 * L2_BYTES // max(set_bytes, 1)) + 1)
    n_sets = min(n_sets, 64)
    sets = [make_set() for _ in range(n_sets)]
    for s in sets[:3]:
        run(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(sets[i % n_sets])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def h(*shape):
    return torch.randn(*shape, device=DEV, dtype=torch.float16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("filters", nargs="*")
    args = ap.parse_args()
    run_cases(args.frames, args.filters)


def run_cases(F_, filters=(), quiet=False):
    """Time every case whose name contains one of `filters` (all if empty); returns the records."""
    H, W = 72, 128
    levels = [(H * W, 320, 5), (H * W // 4, 640, 10), (H * W // 16, 1280, 20), (H * W // 64, 1280, 20)]
    peak, peak_src = peak_gbs()
    cases = []

    for li, (HW, C, heads) in enumerate(levels):
        M = F_ * HW

        def gn_case(fps, C1=C, C2=0, HW=HW, M=M):
            Ct = C1 + C2
            ws = torch.zeros(native.groupnorm_workspace_bytes(F_, HW) // 4 + 1, dtype=torch.float32, device=DEV)

            def mk():
                return (h(M, C1), h(M, C2) if C2 else None, h(Ct), h(Ct), torch.empty(M, Ct, device=DEV, dtype=torch.float16))

            def run(s):
                native.groupnorm_silu(s[4], s[0], s[2], s[3], n_img=F_, HW=HW, eps=1e-5, silu=True, x2=s[1],
                                      frames_per_stat=fps, workspace=ws)
            return mk, run, 2 * M * Ct * 2, 2 * M * Ct * 2   # set bytes, algorithmic bytes (read + write)

        cases.append((f"groupnorm_silu L{li} C={C} per-image", *gn_case(1), "2 launches; reads x twice (stats, apply)"))
        cases.append((f"groupnorm_silu L{li} C={C} per-video (temporal)", *gn_case(F_), "2 launches; reads x twice"))
        if li < 3:
            cases.append((f"groupnorm_silu L{li} cat {C}+{C}", *gn_case(1, C, C), "skip concat fused"))

        def ln_case(M=M, C=C):
            def mk():
                return (h(M, C), h(C), h(C), torch.empty(M, C, device=DEV, dtype=torch.float16))

            def run(s):
                native.layernorm(s[3], s[0], s[1], s[2])
            return mk, run, 2 * M * C * 2, 2 * M * C * 2
        cases.append((f"layernorm L{li} C={C}", *ln_case(), ""))

        def ta_case(M=M, C=C, HW=HW, heads=heads):
            def mk():
                return (h(M, 3 * C), torch.empty(M, C, device=DEV, dtype=torch.float16))

            def run(s):
                native.attn_temporal(s[1], s[0], B=1, F=F_, HW=HW, heads=heads, q_off=0, k_off=C, v_off=2 * C,
                                     scale=0.125)
            return mk, run, 4 * M * C * 2, 4 * M * C * 2
        cases.append((f"attn_temporal L{li} C={C}", *ta_case(), "reads q,k,v once, writes out"))

    # CFG + Euler update and the layout movers at the latent's size
    def euler_case(cfg):
        shape = (1, 4, F_, H, W)
        n = 4 * F_ * H * W

        def mk():
            return (h(*shape), h(F_ * H * W, 4), h(F_ * H * W, 4), h(F_), torch.empty(shape, device=DEV, dtype=torch.float16))

        def run(s):
            native.euler_vpred_step(s[4], s[0], s[1], v_cond=s[2] if cfg else None, gs=s[3] if cfg else None,
                                    v_nhwc=True, c_v=-0.9, c_x=1.5, sigma=2.0, dt=-0.5)
        nb = (4 if cfg else 3) * n * 2
        return mk, run, nb, nb
    cases.append(("euler_vpred (no CFG)", *euler_case(False), "latency-bound at 5.5 MB"))
    cases.append(("euler_vpred (CFG)", *euler_case(True), "latency-bound at 7.4 MB"))

    def up_case(HW=H * W // 4, C=640):
        hh, ww = H // 2, W // 2
        M = F_ * hh * ww

        def mk():
            return (h(M, C), torch.empty(4 * M, C, device=DEV, dtype=torch.float16))

        def run(s):
            native.upsample2x(s[1], s[0], n_img=F_, H=hh, W=ww, Cc=C)
        return mk, run, 5 * M * C * 2, 5 * M * C * 2
    cases.append(("upsample2x 36x64 -> 72x128 C=640", *up_case(), ""))

    def im2col_case(C=320):
        M = F_ * H * W
        Mo = F_ * (H // 2) * (W // 2)

        def mk():
            return (h(M, C), torch.empty(Mo, 9 * C, device=DEV, dtype=torch.float16))

        def run(s):
            native.im2col(s[1], s[0], B=1, F=F_, H=H, W=W, Cc=C, Ho=H // 2, Wo=W // 2, stride=2, taps=native.TAPS_3X3)
        nb = M * C * 2 + Mo * 9 * C * 2
        return mk, run, nb, nb
    cases.append(("im2col stride 2, 72x128 C=320", *im2col_case(), "reads x once (ideal), writes 9/4 x"))

    def frames_case(palette, Fv=25, Hp=576, Wp=1024):
        n = Fv * Hp * Wp

        def mk():      # the permuted view decode_latents returns: [F, 3, H, W] storage seen as [3, F, H, W]
            return (torch.rand(Fv, 3, Hp, Wp, device=DEV) * 2 - 1,)

        def run(s):
            native.frames_to_bytes(s[0].permute(1, 0, 2, 3), rgb=not palette, palette=palette)
        nb = n * 12 + n * (1 if palette else 3)
        return mk, run, nb, nb
    cases.append(("frames_to_bytes 25f 576x1024 fp32 -> RGB bytes", *frames_case(False), "output side of the image -> video run"))
    cases.append(("frames_to_bytes 25f 576x1024 fp32 -> palette indices", *frames_case(True), "fixed 6x7x6 cube, ordered dither"))

    out_path = os.path.join(ROOT, "gpurun_out", "bw_bench.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    records = []
    with open(out_path, "a") as log:
        for name, mk, run, set_bytes, alg_bytes, note in cases:
            if filters and not any(f in name for f in filters):
                continue
            ms = timed(mk, run, set_bytes)
            gbs = alg_bytes / (ms * 1e-3) / 1e9
            rec = {"kernel": name, "frames": F_, "ms": round(ms, 4), "algorithmic_MB": round(alg_bytes / 1e6, 2),
                   "achieved_GBs": round(gbs, 1), "peak_GBs": peak, "frac": round(gbs / peak, 3), "peak_source": peak_src,
                   "variant": os.environ.get("SVDPP_TA_VARIANT", ""), "note": note}
            line = json.dumps(rec)
            records.append(rec)
            if not quiet:
                print(line, flush=True)
            log.write(line + "\n")
            torch.cuda.empty_cache()
    return records


if __name__ == "__main__":
    main()
