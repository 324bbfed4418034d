"""Benchmark mode: stream samples through the step pipeline and print the reference's ``BENCHMARK_JSON=`` line.

Flags, timing method and result keys follow reference ``src/modes/benchmark.py:29-313``: warm-up + measured
samples, per-sample completion times taken on the last rank after a device synchronise, steady throughput =
measured samples / sum of their intervals, first-sample (pipeline fill) time reported separately, peak memory per
rank.  ``--model svd`` builds the native SVD UNet (``--model-id`` = a local diffusers-layout directory or
``random-init[:seed]``; the reference's hub id is mapped to ``random-init`` because there is no network here).
Extensions: ``--allow-uneven`` (25 steps on 8 ranks), ``--schedule ring`` (rotating stage placement, see
``PipelineStage.run_many_ring``).  ``--fsdp`` (the reference's sharded-UNet experiment) is outside this path.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import time

import torch
import torch.distributed as dist

from ..distributed.backend import resolve_backend
from ..distributed.setup import finalize_distributed, init_distributed
from ..pipeline.pipeline import LatentSpec, PipelineConfig, PipelineStage
from ..pipeline.step_assignment import stage_sizes
from ._common import setup_logging

LOGGER = logging.getLogger(__name__)
HUB_ID = "stabilityai/stable-video-diffusion-img2vid-xt"


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Pipeline parallel throughput benchmark")
    p.add_argument("--total-steps", type=int, default=28)
    p.add_argument("--num-samples", type=int, default=10)
    p.add_argument("--latent-channels", type=int, default=4)
    p.add_argument("--latent-frames", type=int, default=14)
    p.add_argument("--latent-height", type=int, default=40)
    p.add_argument("--latent-width", type=int, default=72)
    p.add_argument("--hidden-channels", type=int, default=64)
    p.add_argument("--warmup-samples", type=int, default=2)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--log-level", type=str, default="INFO")
    p.add_argument("--model", type=str, default="dummy", choices=["dummy", "svd"],
                   help="Model to benchmark: dummy (DummyUNet) or svd (native SVD UNet)")
    p.add_argument("--model-id", type=str, default=HUB_ID)
    p.add_argument("--backend", type=str, default="auto", choices=["auto", "gloo", "nccl"])
    p.add_argument("--init-method", type=str, default=None)
    p.add_argument("--guidance-scale", type=float, default=None)
    p.add_argument("--fsdp", action="store_true", help="not supported by this build (outside the step-pipeline path)")
    p.add_argument("--device", type=str, default="cuda", help="extension: 'cpu' runs the dummy model under gloo")
    p.add_argument("--allow-uneven", action="store_true")
    p.add_argument("--schedule", default="fixed", choices=["fixed", "ring"])
    p.add_argument("--transport", default="nccl", choices=["nccl", "peer"],
                   help="extension: 'peer' = peer-mapped receive slots + flags instead of dist.send/recv (CUDA only)")
    return p


def _build_model(args, device):
    if args.model == "dummy":
        from ..models.dummy_unet import DummyUNet
        return DummyUNet(channels=args.latent_channels, hidden_channels=args.hidden_channels).to(device), 1.0, torch.float32
    from ..models.svd_unet import StableVideoUNet
    model_id = "random-init" if args.model_id == HUB_ID and not os.path.isdir(args.model_id) else args.model_id
    model = StableVideoUNet.from_pretrained(model_id=model_id,
                                            timesteps=StableVideoUNet._default_timestep_schedule(args.total_steps),
                                            torch_dtype=torch.float16, device=device)
    model.enable_memory_optimizations()
    model.set_dummy_conditioning(batch_size=1, num_frames=args.latent_frames, height=args.latent_height,
                                 width=args.latent_width, device=device, guidance_scale=args.guidance_scale)
    model.use_cuda_graph = True
    return model, model.init_noise_sigma, torch.float16


def main(argv=None) -> dict | None:
    args = build_parser().parse_args(argv)
    setup_logging(args.log_level)
    if args.fsdp:
        raise SystemExit("--fsdp (sharded-UNet experiment) is outside the step-pipeline path this build covers")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    on_gpu = args.device != "cpu"
    if not on_gpu and args.model == "svd":
        raise SystemExit("--model svd needs a GPU: the native SVD UNet has no CPU path")
    backend = resolve_backend(None if args.backend == "auto" else args.backend, simulator=not on_gpu)
    device = torch.device(f"cuda:{local_rank}") if on_gpu else torch.device("cpu")
    if on_gpu:
        torch.cuda.set_device(device)
    init_distributed(backend=backend, rank=rank, world_size=world, init_method=args.init_method)

    def sync():
        if on_gpu:
            torch.cuda.synchronize(device)

    model, noise_sigma, dtype = _build_model(args, device)
    shape = torch.Size((1, args.latent_channels, args.latent_frames, args.latent_height, args.latent_width))
    total = args.warmup_samples + args.num_samples
    cfg = PipelineConfig(total_steps=args.total_steps, world_size=world, rank=rank,
                         timesteps=list(range(args.total_steps - 1, -1, -1)),   # as the reference (benchmark.py:178)
                         latent_spec=LatentSpec(shape=shape, dtype=dtype, device=device), allow_uneven=args.allow_uneven)
    stage = PipelineStage(model=model, config=cfg, transport=args.transport if on_gpu else "nccl")

    def supplier(idx: int) -> torch.Tensor:
        torch.manual_seed(args.seed + idx)
        return torch.randn(shape, device=device, dtype=dtype) * noise_sigma

    if on_gpu:
        torch.cuda.reset_peak_memory_stats(device)
    sync()
    if world > 1:
        dist.barrier()
    ends = []
    start = time.perf_counter()
    with torch.no_grad():
        if args.schedule == "ring" and world > 1:
            # rotating placement: every rank finishes one video per batch of `world`; completion times are per batch
            for b in range(0, total, world):
                n = min(world, total - b)
                stage.run_many_ring(n, input_supplier=lambda i, b=b: supplier(b + i))
                sync()
                if world > 1:
                    dist.barrier()
                t = time.perf_counter()
                ends.extend([t] * n)
        else:
            for idx in range(total):
                stage._process_single_latent(supplier(idx) if rank == 0 else None, sample_idx=idx)
                if rank == world - 1:
                    sync()
                    ends.append(time.perf_counter())
    sync()
    peak = torch.cuda.max_memory_allocated(device) if on_gpu else 0
    peaks = [peak]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, peak)
        peaks = gathered
    results = None
    if rank == world - 1:
        per_sample = [e - (start if i == 0 else ends[i - 1]) for i, e in enumerate(ends)]
        if args.schedule == "ring" and world > 1:   # a batch's samples complete together: spread its interval
            per_sample, prev = [], start
            for b in range(0, total, world):
                n = min(world, total - b)
                per_sample.extend([(ends[b] - prev) / n] * n)
                prev = ends[b]
        measured = per_sample[args.warmup_samples:]
        mtot = sum(measured)
        results = {
            "world_size": world, "total_steps": args.total_steps,
            "steps_per_gpu": args.total_steps // world if args.total_steps % world == 0 else stage_sizes(args.total_steps, world),
            "model": args.model, "fsdp": False, "num_samples_measured": args.num_samples,
            "warmup_samples": args.warmup_samples, "latent_shape": list(shape),
            "first_sample_time_s": round(per_sample[0], 4) if per_sample else 0.0,
            "avg_sample_time_s": round(mtot / len(measured), 4) if measured else 0.0,
            "throughput_samples_per_s": round(len(measured) / mtot, 4) if mtot > 0 else 0.0,
            "per_sample_times_ms": [round(t * 1000, 2) for t in per_sample],
            "peak_memory_gb_per_rank": [round(p / 1e9, 3) for p in peaks],
            "max_peak_memory_gb": round(max(peaks) / 1e9, 3),
            "schedule": args.schedule,
        }
        LOGGER.info("BENCHMARK RESULTS (Pipeline mode): GPUs %d | model %s | first sample %.2f s | steady %.4f s | "
                    "%.4f samples/s", world, args.model, results["first_sample_time_s"], results["avg_sample_time_s"],
                    results["throughput_samples_per_s"])
        print(f"BENCHMARK_JSON={json.dumps(results)}", flush=True)
    finalize_distributed()
    return results


if __name__ == "__main__":
    main()
