"""bench.py — SVD-XT step-pipeline throughput on B200 (contract: see the task statement / DESIGN.md).

Default arm (ours):
  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun, one rank per GPU)
Workload: SVD-XT, 25 frames, 576x1024 (latent 72x128), 25 Euler steps, random-init fp16 UNet
(1 524 623 082 parameters), dummy image conditioning, no CFG unless --guidance-scale is given.
One bench "step" = one video through all 25 denoising steps.  With N GPUs the 25 steps are split into
N pipeline stages (uneven split allowed) and K*N videos are streamed, so per-GPU work is fixed
("weak").  value = videos/min over the whole timed region (pipeline fill and drain included).
Stage placement for N>1 is --schedule ring by default (stage s of video v on rank (v+s)%N, see
PipelineStage.run_many_ring); the reference's fixed placement is timed as well and reported under
"linear_pipeline".

Reference arm:
  python bench.py --impl reference ...   times the CPU restatement of the reference's step (oracle/) on
  the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SVD-XT 25f 576x1024 videos/min (25 Euler steps, step-pipeline)"
UNIT = "videos/min"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2, help="timed bench steps (videos per GPU)")
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--frames", type=int, default=25)
    p.add_argument("--latent-height", type=int, default=72)
    p.add_argument("--latent-width", type=int, default=128)
    p.add_argument("--denoise-steps", type=int, default=25)
    p.add_argument("--guidance-scale", type=float, default=None)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--no-graph", action="store_true",
                   help="enqueue every launch eagerly instead of replaying one CUDA graph per denoising step (within "
                        "+-1 %%: the host runs ~740 launches per step ahead of the GPU either way)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-sample-frames", type=int, default=2)
    p.add_argument("--schedule", default="ring", choices=["ring", "linear"],
                   help="N>1: 'ring' rotates the stage->rank placement per video (no fill/drain bubble, no stage "
                        "imbalance); 'linear' is the reference's fixed placement (stage s on rank s)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_sample(frames: int, H: int, W: int, total_steps: int, repeats: int = 1):
    """One denoising step of the oracle (torch restatement of the reference's wrapper + UNet) on the
    host cores, fp32, at `frames` frames of the full 72x128 latent.  Returns (seconds/step, threads)."""
    import torch
    from oracle.svd_step import OracleStep, dummy_conditioning
    from oracle.unet_torch import UNetSpatioTemporalConditionModel
    # all host cores: torchrun exports OMP_NUM_THREADS=1 for N > 1, which would time a single-thread baseline
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.device("meta"):
        unet = UNetSpatioTemporalConditionModel()
    unet = unet.to_empty(device="cpu")
    g = torch.Generator().manual_seed(0)
    for mod in unet.modules():
        if isinstance(mod, (torch.nn.GroupNorm, torch.nn.LayerNorm)):
            mod.weight.data.fill_(1.0)
            mod.bias.data.zero_()
    for name, prm in unet.named_parameters():
        if prm.dim() >= 2:
            fan_in = prm[0].numel()
            prm.data.uniform_(-fan_in ** -0.5, fan_in ** -0.5, generator=g)
        elif "norm" not in name:
            prm.data.zero_()
    for name, prm in unet.named_parameters():
        if name.endswith("mix_factor"):
            prm.data.fill_(0.5)
    unet.eval()
    step = OracleStep(unet, total_steps, dtype=torch.float32)
    torch.manual_seed(1)
    cond = dummy_conditioning(1, frames, H, W, torch.device("cpu"), torch.float32)
    lat = torch.randn(1, 4, frames, H, W) * step.init_noise_sigma
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = step(lat, 0, cond)
        times.append(time.perf_counter() - t0)
    assert bool(torch.isfinite(out).all())
    return min(times), torch.get_num_threads()


def videos_per_min_from_sample(sec_per_step_sample: float, sample_frames: int, frames: int, total_steps: int) -> float:
    # UNet cost is linear in frames to within 0.1 % (only temporal attention is not)
    sec_video = sec_per_step_sample * (frames / sample_frames) * total_steps
    return 60.0 / sec_video


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    H, W = args.latent_height, args.latent_width
    sf = args.cpu_sample_frames
    times = []
    n = max(1, args.steps)
    for i in range(max(0, min(args.warmup, 1)) + n):
        t, threads = cpu_oracle_sample(sf, H, W, args.denoise_steps)
        if i >= max(0, min(args.warmup, 1)):
            times.append(t)
    sec = statistics.mean(times)
    v = videos_per_min_from_sample(sec, sf, args.frames, args.denoise_steps)
    sample = (f"1 denoising step (UNet forward + Euler update) of the oracle, fp32 on CPU, {sf} of {args.frames} "
              f"frames at the full {H}x{W} latent; scaled by frames/{sf} x {args.denoise_steps} steps")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": 60000.0 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SVD-XT {args.frames}f 576x1024 latent {H}x{W}, {args.denoise_steps} steps, no CFG",
                   "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "sec_per_sample_step": sec},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_dummy_simulator(total_steps: int = 25):
    """The reference's CPU simulator workload (BASELINE config 1): DummyUNet(channels=4), latent [1,4,14,64,64]
    fp32, `total_steps` steps on the host cores in one process (torch CPU kernels, as the reference runs it)."""
    import torch
    from vdpp_b200.models import DummyUNet
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(1234)
    model = DummyUNet(channels=4).eval()
    torch.manual_seed(42)
    lat = torch.randn(1, 4, 14, 64, 64)
    with torch.no_grad():
        lat = model(lat, 0)          # warm-up
        t0 = time.perf_counter()
        for step in range(total_steps):
            lat = model(lat, step)
        sec = time.perf_counter() - t0
    return {"workload": f"DummyUNet(channels=4) latent 1x4x14x64x64 fp32, {total_steps} steps, 1 process",
            "ms_per_step": 1000.0 * sec / total_steps, "videos_per_min": 60.0 / sec, "cores": torch.get_num_threads(),
            "finite": bool(torch.isfinite(lat).all())}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ our arm
def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import vdpp_b200  # noqa: F401
    from vdpp_b200 import native
    from vdpp_b200.models import StableVideoUNet
    from vdpp_b200.models.native_unet import flops_per_forward
    from vdpp_b200.pipeline import LatentSpec, PipelineConfig, PipelineStage, stage_sizes

    # stdout carries exactly ONE JSON line (the contract).  NCCL prints its banner ("NCCL version ...") with printf on
    # file descriptor 1, so fd 1 is pointed at stderr for the whole run and the line goes out through a saved copy.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    F_, H, W, T = args.frames, args.latent_height, args.latent_width, args.denoise_steps
    K, Wm = args.steps, max(args.warmup, 3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass

    # ---- model: random-init SVD-XT UNet behind the reference's wrapper API
    t_build = time.time()
    model = StableVideoUNet.from_pretrained("random-init:0", timesteps=StableVideoUNet._default_timestep_schedule(T),
                                            device=dev)
    torch.manual_seed(args.seed + 1)
    model.set_dummy_conditioning(batch_size=1, num_frames=F_, height=H, width=W, device=dev,
                                 guidance_scale=args.guidance_scale)
    model.use_cuda_graph = not args.no_graph
    t_build = time.time() - t_build
    shape = torch.Size((1, 4, F_, H, W))
    spec = LatentSpec(shape=shape, dtype=torch.float16, device=dev)
    cfg = PipelineConfig(total_steps=T, world_size=world, rank=rank, timesteps=list(range(T)), latent_spec=spec,
                         allow_uneven=True)
    stage = PipelineStage(model=model, config=cfg)
    n_videos = K * world
    last = rank == world - 1

    def make_inputs(n, base):
        out = []
        for i in range(n):
            g = torch.Generator(device=dev).manual_seed(args.seed + base + i)
            out.append(torch.randn(shape, device=dev, generator=g, dtype=torch.float32).half() * model.init_noise_sigma)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ring = world > 1 and args.schedule == "ring"

    def run_stream(n, supplier):
        """Stream n videos; returns the latents that finished on this rank."""
        if ring:
            return [o for _, o in stage.run_many_ring(n, input_supplier=supplier)]
        return stage.run_many(n, input_supplier=supplier if rank == 0 else None) or []

    # ---- warm-up: first step eager, then one CUDA graph per step index is captured, then replays
    n_warm = Wm * world if ring else Wm
    warm_in = {i: t for i, t in zip(range(n_warm), make_inputs(n_warm, 1000))} if (ring or rank == 0) else {}
    run_stream(n_warm, lambda i: warm_in[i])
    del warm_in
    barrier()

    # ---- timed region 1: inputs resident in HBM
    inputs = make_inputs(n_videos, 0) if (ring or rank == 0) else None
    sampler = ClockSampler(local_rank)
    launches0 = native.LAUNCHES
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = run_stream(n_videos, lambda i: inputs[i])
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = torch.tensor([native.LAUNCHES - launches0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches)
    finite = all(bool(torch.isfinite(o).all()) for o in outs)

    # ---- timed region 2: end to end through the public API with host buffers
    need_in = ring or rank == 0
    host_in = [torch.randn(shape, dtype=torch.float32).half().mul_(model.init_noise_sigma).pin_memory()
               if (need_in and (not ring or i % world == rank)) else None for i in range(n_videos)]
    host_out = torch.empty(shape, dtype=torch.float16).pin_memory() if (ring or last) else None

    def supply_from_host(i):
        return host_in[i].to(dev, non_blocking=True)

    barrier()
    t0 = time.perf_counter()
    if ring:
        for _, lat in stage.run_many_ring(n_videos, input_supplier=supply_from_host):
            host_out.copy_(lat, non_blocking=True)
        torch.cuda.synchronize()
    else:
        for i in range(n_videos):
            if rank == 0:
                lat = supply_from_host(i)
            else:
                work, buf = stage._post_recv()
                work.wait()
                lat = buf if stage.step_range.count else buf.clone()
            lat = stage._run_local_steps(lat)
            if last:
                host_out.copy_(lat, non_blocking=True)
                torch.cuda.synchronize()
            else:
                stage._send_latent(lat, blocking=False)
                stage._drain_send()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    # ---- N>1: the reference's fixed stage placement on the same stream, for comparison
    linear_value = None
    if ring:
        n_lin = 2 * world
        lin_in = make_inputs(n_lin, 500) if rank == 0 else None
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        stage.run_many(n_lin, input_supplier=(lambda i: lin_in[i]) if rank == 0 else None)
        l1.record()
        barrier()
        lms = torch.tensor([l0.elapsed_time(l1)], device=dev)
        dist.all_reduce(lms, op=dist.ReduceOp.MAX)
        linear_value = n_lin / (float(lms.item()) / 1000.0) * 60.0
    bytes_lat = shape.numel() * 2

    if rank == 0:
        value = n_videos / (ms_total / 1000.0) * 60.0
        e2e_value = n_videos / e2e_s * 60.0
        fl = flops_per_forward(model.unet.cfg, 2 if args.guidance_scale and args.guidance_scale > 1 else 1, F_, H, W)
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 (fp32 accumulate)", "data": "synthetic",
            "config": {
                "workload": f"SVD-XT {F_}f 576x1024 (latent {H}x{W}), {T} Euler steps, random-init fp16 UNet "
                            f"1524623082 params, dummy conditioning, "
                            f"{'CFG %.1f batch 2' % args.guidance_scale if args.guidance_scale else 'no CFG'}",
                "videos_timed": n_videos, "stage_sizes": stage_sizes(T, world),
                "parallelism": (f"step-pipeline x{world}, {'rotating (ring)' if ring else 'fixed (reference)'} stage "
                                f"placement") if world > 1 else "single GPU",
                "cuda_graph": model.use_cuda_graph,
                "l2": "per-step working set (3 GB weights + >10 GB activations) exceeds the 126 MB L2; no flush needed",
            },
            "unet_step_ms": ms_total / (n_videos * T / world) if world == 1 else None,
            "unet_tflop_per_forward": fl["total"] / 1e12,
            "gpu_launches": int(launches.item()),
            "clocks": clocks,
            "output_finite": finite,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_lat, "d2h_bytes_per_step": bytes_lat},
            "model_build_s": round(t_build, 1),
        }
        if linear_value is not None:
            result["linear_pipeline"] = {"value": linear_value, "unit": UNIT, "videos_timed": 2 * world,
                                         "note": "reference placement (stage s on rank s), fill/drain included"}
    # ---- roofline of the dominant kernel (tcgen05 GEMM/conv family), instrumented eager pass, rank 0
    if rank == 0:
        model.use_cuda_graph = False
        native.PROFILE = []
        x = make_inputs(1, 0)[0]
        torch.cuda.synchronize()
        model(x, 0)
        torch.cuda.synchronize()
        prof, native.PROFILE = native.PROFILE, None
        agg = {}
        for kind, flops, _shape, a, b in prof:
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += flops
            d[2] += 1
        gemm_ms = sum(agg[k][0] for k in ("conv", "linear", "geglu") if k in agg)
        gemm_flops = fl["conv3x3"] + fl["conv_up"] + fl["conv_t"] + fl["conv1x1"] + fl["linear"] + fl["geglu_ff"]
        n_gemm = sum(agg[k][2] for k in ("conv", "linear", "geglu") if k in agg)
        peak = peaks.get("bf16_tflops_sustained")
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        if peak is None:
            peak, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of the same kernel family, from the committed ncu pass (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_dram_traffic.json")))
            if F_ == 25 and (H, W) == (72, 128) and not args.guidance_scale:
                traffic, traffic_src = tj["traffic_bytes_per_launch"], "profiles/r1_gemm_dram_traffic.json (" + tj["source"] + ")"
        except Exception:  # noqa: BLE001
            pass
        result["roofline"] = {
            "kernel": "gemm_tc_kernel<160|256|128,*> (tcgen05 GEMM + implicit-GEMM conv, all launches of one UNet forward)",
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "launches": n_gemm, "avg_launch_ms": gemm_ms / max(n_gemm, 1),
            "algorithmic_tflop_per_forward": gemm_flops / 1e12,
        }
        att = agg.get("attn_spatial")
        if att:
            result["roofline_attention"] = {
                "kernel": "attn_spatial2_tc_kernel / attn_spatial_tc_kernel (tcgen05 FMHA)", "bound": "tensor",
                "achieved": fl["attn_spatial"] / (att[0] / 1000.0) / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": fl["attn_spatial"] / (att[0] / 1000.0) / 1e12 / peak, "launches": att[2], "ms": att[0],
                # head_dim 64: one MUFU.EX2 per score (16/clk/SM, tools/ubench) against 256 tensor FLOP per score
                "mufu_bound_tflops": 16 * 148 * (clocks.get("sm_mhz") or 1965.0) * 1e6 * 256 / 1e12,
                "note": "bound by the MUFU pipe before the tensor pipe; mufu_bound_tflops at the SM clock seen in the timed region"}
        result["kernel_ms_per_forward"] = {k: round(v[0], 3) for k, v in agg.items()}
        if F_ == 25 and (H, W) == (72, 128):
            try:   # achieved HBM GB/s of the bandwidth kernels at this workload's level-0 shapes (tools/bw_bench.py)
                from tools.bw_bench import run_cases
                del x
                torch.cuda.empty_cache()
                result["roofline_bandwidth"] = run_cases(F_, ["L0 C=320 per-image", "L0 C=320 per-video", "layernorm L0",
                                                              "attn_temporal L0", "euler_vpred (no CFG)"], quiet=True)
            except Exception as e:  # noqa: BLE001
                result["roofline_bandwidth"] = {"error": f"{type(e).__name__}: {e}"}
        result["whole_step_tflops"] = (fl["total"] / 1e12) / (ms_total / 1000.0 / (n_videos * T / world)) if world == 1 else None

    # ---- CPU baseline (oracle port on the host cores), rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sec, threads = cpu_oracle_sample(args.cpu_sample_frames, H, W, T)
            v = videos_per_min_from_sample(sec, args.cpu_sample_frames, F_, T)
            result["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"1 denoising step of the oracle (torch fp32, CPU) at {args.cpu_sample_frames} of {F_} frames, "
                          f"full {H}x{W} latent, scaled by frames x {T} steps; {sec:.2f} s measured"}
            result["cpu_baseline"]["reference_simulator"] = cpu_dummy_simulator(T)
        except Exception as e:  # noqa: BLE001
            result["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                      "sample": f"failed: {type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(result), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
