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

For N>1 the line also carries, for BOTH placements, the reference benchmark mode's own metrics (2 warm-up + 14
measured videos, per-video completion times on the finishing rank: steady videos/min and first-video latency,
reference src/modes/benchmark.py:256-267) and "multi_gpu_bit_equal": videos that went through the pipeline (incl.
ones whose receive slot is a reused one) are recomputed on one GPU and must be torch.equal.
At N=1 the line carries "library_baseline": the torch restatement of the same UNet with the same weights run with
torch's library kernels (cuDNN / cuBLAS / SDPA, eager fp16) - BASELINE.md section 4.

Reference arm:
  python bench.py --impl reference ...   times the CPU restatement of the reference's step (oracle/) on
  the host cores on a bounded sample of the same workload (the reference's own SVD code cannot run here:
  its UNet lives in diffusers, which is not installed; its own CPU simulator code IS timed, see
  cpu_baseline.reference_simulator).
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
    p.add_argument("--transport", default="peer", choices=["nccl", "peer"],
                   help="N>1 latent handoff: 'peer' (default) = the next stage's peer-mapped receive slot written by the "
                        "producer's Euler kernel + flag (distributed/handoff.py; falls back to nccl on every rank if the "
                        "symmetric-memory rendezvous fails anywhere), 'nccl' = dist.send/recv waited on the stream")
    p.add_argument("--force-graph", action="store_true", help="N=1: skip the eager-vs-graph measurement, use graphs")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-library-baseline", action="store_true")
    p.add_argument("--cpu-sample-frames", type=int, default=6,
                   help="frames of the CPU baseline's bounded sample: 6 of 25 = 12.3 s of CPU work per step on the box's 16 "
                        "cores (2 frames: 3.9 s, a 4 %% better CPU number per frame)")
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
                         "sec_per_sample_step": sec, "extrapolated": True,
                         "extrapolation": f"x {args.frames}/{sf} frames x {args.denoise_steps} steps"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_dummy_simulator(total_steps: int = 25):
    """The reference's CPU simulator workload (BASELINE config 1: DummyUNet(channels=4), latent [1,4,14,64,64] fp32) on
    the host cores, run with the REFERENCE'S OWN PipelineStage + DummyUNet when an installed copy of the reference is
    reachable (baseline/_ref or $VDPP_REFERENCE_ROOT; tools/reference_simulator.py), else with this repository's
    mirror of those classes ("kind": "mirror").  25 steps x 1 rank, and 28 steps x 4 ranks over gloo (25 x 4 is
    rejected by the reference's divisibility rule)."""
    tool = os.path.join(ROOT, "tools", "reference_simulator.py")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT",
                                                           "OMP_NUM_THREADS", "PYTHONPATH")}
    runs = {}
    for name, cmd in (
            (f"{total_steps}x1", [sys.executable, tool, "--total-steps", str(total_steps), "--samples", "4", "--warmup", "2"]),
            ("28x4", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4",
                      "--master-addr", "127.0.0.1", "--master-port", "29688", tool, "--total-steps", "28",
                      "--samples", "6", "--warmup", "2"])):
        try:
            r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("REFERENCE_SIMULATOR_JSON=")]
            runs[name] = json.loads(line[-1].split("=", 1)[1]) if line else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as e:  # noqa: BLE001
            runs[name] = {"error": f"{type(e).__name__}: {e}"}
    first = runs.get(f"{total_steps}x1", {})
    if "error" not in first:
        return {"kind": "reference", "runs": runs, "ms_per_step": first["ms_per_step_per_rank"],
                "videos_per_min": first["videos_per_min"], "cores": first["threads_per_rank"]}
    # no reference tree: time this repository's mirror of the same classes in-process
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
    return {"kind": "mirror", "why": first.get("error"),
            "workload": f"DummyUNet(channels=4) latent 1x4x14x64x64 fp32, {total_steps} steps, 1 process",
            "ms_per_step": 1000.0 * sec / total_steps, "videos_per_min": 60.0 / sec, "cores": torch.get_num_threads(),
            "finite": bool(torch.isfinite(lat).all())}


def library_baseline(model, dev, F_, H, W, T, guidance, seed):
    """the torch restatement of the UNet (oracle/unet_torch.py) with the SAME weights, run by
    torch's library kernels (cuDNN conv, cuBLASLt GEMM, SDPA attention, ATen norms) in eager fp16 behind the oracle's
    restatement of the reference wrapper - 'the reference's own torch fp16 pipeline'.  A measured baseline beside the
    native number; never on the product path."""
    import torch
    from oracle.svd_step import Conditioning, OracleStep
    from oracle.unet_torch import UNetSpatioTemporalConditionModel
    from vdpp_b200.models.svd_weights import random_state_dict
    with torch.device("meta"):
        lib = UNetSpatioTemporalConditionModel(norm_eps=model.unet.cfg["norm_eps"])
    lib = lib.to_empty(device=dev).half().eval()
    lib.load_state_dict(random_state_dict(None, seed=0, device=dev), strict=True)
    cond = Conditioning(model._image_embeddings, model._image_latents, dtype=torch.float16, guidance_scale=guidance,
                        num_frames=F_)
    ostep = OracleStep(lib, T)
    g = torch.Generator(device=dev).manual_seed(seed)
    x0 = torch.randn((1, 4, F_, H, W), device=dev, generator=g, dtype=torch.float32).half() * model.init_noise_sigma
    x = ostep(x0, 0, cond)                                    # warm-up (cuDNN / cuBLASLt heuristics)
    native_1 = model(x0, 0)
    n_steps = 3
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in range(n_steps):
        x = ostep(x, s_ + 1, cond)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_steps
    lib_1 = ostep(x0, 0, cond)
    d = (native_1.float() - lib_1.float()).abs()
    out = {"what": "torch eager fp16 (cuDNN/cuBLASLt/SDPA/ATen), same weights, same wrapper arithmetic",
           "ms_per_step": ms, "videos_per_min": 60000.0 / (ms * T), "steps_timed": n_steps,
           "first_step_max_abs_native_vs_library": d.max().item(), "latent_absmax": lib_1.float().abs().max().item(),
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    del lib, ostep, cond
    torch.cuda.empty_cache()
    return out


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
    stage = PipelineStage(model=model, config=cfg, transport=args.transport if world > 1 else "nccl")
    transport_note = None
    if world > 1 and stage.transport == "peer":
        # the peer-mapped slots need a symmetric-memory rendezvous; if it fails on ANY rank, every rank uses NCCL
        try:
            stage._peer_handoff()
            ok = 1
        except Exception as e:  # noqa: BLE001
            ok, transport_note = 0, f"peer handoff unavailable on rank {rank}: {type(e).__name__}: {e}"
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            stage.transport, stage._peer = "nccl", None
            transport_note = transport_note or "peer handoff unavailable on another rank"
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

    def run_stream(n, supplier, use_ring=None):
        """Stream n videos; returns {video index: final latent} for the videos that finished on this rank."""
        use_ring = ring if use_ring is None else use_ring
        if use_ring:
            return dict(stage.run_many_ring(n, input_supplier=supplier))
        outs = stage.run_many(n, input_supplier=supplier if rank == 0 else None) or []
        return dict(enumerate(outs))

    def solo(x):
        """All T steps of one video on this GPU alone (no pipeline): the bit-equality reference."""
        for s_ in cfg.timesteps:
            x = model(x, s_)
        return x

    # ---- warm-up: first step eager, then one CUDA graph per step index is captured, then replays
    n_warm = Wm * world if ring else Wm
    warm_in = {i: t for i, t in zip(range(n_warm), make_inputs(n_warm, 1000))} if (ring or rank == 0) else {}
    run_stream(n_warm, lambda i: warm_in[i])
    barrier()

    # ---- N=1: eager enqueue vs one CUDA graph per step, measured here, the faster one is used for the timed regions
    launch_mode = None
    if world == 1 and not args.no_graph and not args.force_graph:
        x = warm_in[0]
        per_mode = {}
        for mode in (False, True):
            model.use_cuda_graph = mode
            for s_ in range(min(T, 3)):            # captures the graphs of these steps on the first pass
                model(x, s_)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(4):
                for s_ in range(min(T, 3)):
                    model(x, s_)
            g1.record()
            torch.cuda.synchronize()
            per_mode["graph" if mode else "eager"] = g0.elapsed_time(g1) / (4 * min(T, 3))
        model.use_cuda_graph = per_mode["graph"] <= per_mode["eager"]
        launch_mode = {"step_ms_eager": per_mode["eager"], "step_ms_graph": per_mode["graph"],
                       "chosen": "graph" if model.use_cuda_graph else "eager"}
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
    finite = all(bool(torch.isfinite(o).all()) for o in outs.values())

    # ---- N>1: a video that went through the pipeline must equal the same video computed on one GPU, bit for bit
    # (NCCL handoffs, CUDA-graph replays, receive slots reused two hops later).  Every rank re-runs up to two of the
    # videos that finished on it - in ring placement video v ends on rank (v + N - 1) % N, so videos of the first AND
    # of later batches are covered; in fixed placement everything ends on the last rank.
    bit_equal = None
    if world > 1:
        mine = sorted(outs)
        picks = ([mine[0], mine[-1]] if len(mine) > 1 else mine)
        ok = 1
        for v in picks:
            ref_in = inputs[v] if inputs is not None else make_inputs(v + 1, 0)[v]
            ok &= int(torch.equal(solo(ref_in), outs[v]))
        flag = torch.tensor([ok], device=dev)
        n_checked = torch.tensor([len(picks)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.all_reduce(n_checked)
        bit_equal = {"ring" if ring else "fixed": bool(flag.item()), "videos_checked": int(n_checked.item())}
    del outs

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
            lat = stage._process_single_latent(supply_from_host(i) if rank == 0 else None, sample_idx=i)
            if last:
                host_out.copy_(lat, non_blocking=True)
                torch.cuda.synchronize()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    # ---- N>1: both placements measured the way the reference's benchmark mode does (src/modes/benchmark.py:229-267,
    # SURVEY 8d cfg-3): 2 warm-up + 14 measured videos, completion times taken after a device synchronise on the rank a
    # video finishes on; steady = measured videos / sum of their completion intervals; first-video time = pipeline fill.
    placements = None
    if world > 1:
        n_warm_s, n_meas = 2, 14
        placements = {}
        for name in ("fixed", "ring"):
            tot = n_warm_s + n_meas
            if name == "ring":      # whole batches of `world` videos: completion times are per batch
                tot = -(-tot // world) * world
            s_in = make_inputs(tot, 2000) if (name == "ring" or rank == 0) else None
            ends = []
            barrier()
            t_start = time.perf_counter()
            if name == "fixed":
                for i in range(tot):
                    out_i = stage._process_single_latent(s_in[i] if rank == 0 else None, sample_idx=i)
                    if last:
                        torch.cuda.synchronize()
                        ends.append(time.perf_counter())
                        if i in (0, 2, tot - 1):          # slots 0 / 0 again (first reuse) / last: bit-equality below
                            placements.setdefault("_keep", {})[i] = out_i.clone()
            else:
                for b in range(0, tot, world):
                    stage.run_many_ring(world, input_supplier=lambda i, b=b: s_in[b + i])
                    torch.cuda.synchronize()
                    dist.barrier()
                    ends.extend([time.perf_counter()] * world)
            barrier()
            if name == "fixed":
                per = [e - (t_start if i == 0 else ends[i - 1]) for i, e in enumerate(ends)] if last else []
            else:
                per, prev = [], t_start
                for b in range(0, tot, world):
                    per.extend([(ends[b] - prev) / world] * world)
                    prev = ends[b]
            src_rank = world - 1 if name == "fixed" else 0
            stats = torch.zeros(3, device=dev, dtype=torch.float64)
            if rank == src_rank:
                meas = per[n_warm_s:n_warm_s + n_meas] if name == "fixed" else per[world:]
                first = per[0] if name == "fixed" else (ends[0] - t_start)
                stats = torch.tensor([len(meas) / sum(meas) * 60.0, first, float(len(meas))], device=dev, dtype=torch.float64)
            dist.broadcast(stats, src=src_rank)
            placements[name] = {"steady_videos_per_min": float(stats[0]), "first_video_s": float(stats[1]),
                                "videos_measured": int(stats[2]), "videos_warmup": n_warm_s if name == "fixed" else world}
            if name == "fixed":
                kept = placements.pop("_keep", {})
                ok = 1
                if last:
                    again = make_inputs(tot, 2000)
                    for i, o in kept.items():
                        ok &= int(torch.equal(solo(again[i]), o))
                    del again
                flag = torch.tensor([ok], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                bit_equal["fixed" if ring else "fixed_steady_run"] = bool(flag.item())
            del s_in
        sizes = stage_sizes(T, world)
        placements["ideal_efficiency_fixed"] = T / (world * max(sizes))
        placements["method"] = ("reference src/modes/benchmark.py:229-267: completion times after a device synchronise, "
                                "steady = measured / sum of intervals; ring placement completes a batch of N videos at once")
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
                                f"placement, handoff via {stage.transport}") if world > 1 else "single GPU",
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
        if launch_mode is not None:
            result["launch_mode"] = launch_mode
        if world > 1:
            result["handoff"] = {"transport": stage.transport, "note": transport_note,
                                 "bytes_per_boundary_per_video": bytes_lat}
        if placements is not None:
            result["placements"] = placements
            result["linear_pipeline"] = {"value": placements["fixed"]["steady_videos_per_min"], "unit": UNIT,
                                         "note": "reference placement (stage s on rank s), steady; see placements"}
        if bit_equal is not None:
            result["multi_gpu_bit_equal"] = all(v for k, v in bit_equal.items() if k != "videos_checked")
            result["multi_gpu_bit_equal_detail"] = bit_equal
    # ---- roofline of the dominant kernel (tcgen05 GEMM/conv family), instrumented eager pass, rank 0
    if rank == 0:
        # per-kernel CUDA events need one ctypes launch per kernel: the Python orchestration of the same launch sequence
        # (bit-identical to csrc/unet.cu's, tests/test_gpu_unet.py) with the same weights and conditioning
        prof_model = StableVideoUNet.from_pretrained("random-init:0", timesteps=StableVideoUNet._default_timestep_schedule(T),
                                                     device=dev, orchestrator="python")
        prof_model.set_conditioning(model._image_embeddings, model._image_latents, guidance_scale=args.guidance_scale,
                                    num_frames=F_)
        x = make_inputs(1, 0)[0]
        prof_model(x, 0)                      # warm (frame-position cache, workspaces)
        torch.cuda.synchronize()
        native.PROFILE = []
        prof_model(x, 0)
        torch.cuda.synchronize()
        prof, native.PROFILE = native.PROFILE, None
        del prof_model
        torch.cuda.empty_cache()
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
            tname = next(n for n in ("r2_gemm_dram_traffic.json", "r1_gemm_dram_traffic.json")
                         if os.path.exists(os.path.join(ROOT, "profiles", n)))
            tj = json.load(open(os.path.join(ROOT, "profiles", tname)))
            if F_ == 25 and (H, W) == (72, 128) and not args.guidance_scale:
                traffic = tj["traffic_bytes_per_launch"]
                traffic_src = f"profiles/{tname} ({tj['source']}; {tj.get('launches', '?')} launches then, {n_gemm} now)"
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
                "kernel": "attn_spatial3_tc_kernel (S >= 1024, ping-pong) / attn_spatial_tc_kernel (tcgen05 FMHA)", "bound": "tensor",
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

    # ---- GPU library baseline (torch eager fp16, same weights), rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            result["library_baseline"] = library_baseline(model, dev, F_, H, W, T, args.guidance_scale, args.seed)
            result["library_baseline"]["native_speedup"] = result["library_baseline"]["ms_per_step"] / result["unet_step_ms"]
        except Exception as e:  # noqa: BLE001
            result["library_baseline"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline (oracle port on the host cores), rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sec, threads = cpu_oracle_sample(args.cpu_sample_frames, H, W, T)
            v = videos_per_min_from_sample(sec, args.cpu_sample_frames, F_, T)
            result["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "extrapolated": True, "extrapolation": f"x {F_}/{args.cpu_sample_frames} frames x {T} steps",
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
