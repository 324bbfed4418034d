"""Data-parallel baseline mode: every rank runs whole videos (all steps) on its own GPU, no latent exchange.

Flags and the ``BENCHMARK_JSON=`` keys follow reference ``src/modes/benchmark_data_parallel.py:28-262``: each rank
warms up locally, the measured samples are split across ranks (``ceil(num_samples / world)`` each, seeds
``seed + sample_idx``), throughput = measured samples / the slowest rank's wall clock, reported by rank 0.  This is
the "replicas" figure the reference compares its step pipeline against.
Extension: ``--device cpu`` runs the dummy model under gloo.
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
from ._common import setup_logging
from .benchmark import HUB_ID, _build_model

LOGGER = logging.getLogger(__name__)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Data parallel throughput benchmark")
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
    p.add_argument("--model", type=str, default="dummy", choices=["dummy", "svd"])
    p.add_argument("--model-id", type=str, default=HUB_ID)
    p.add_argument("--backend", type=str, default="auto", choices=["auto", "gloo", "nccl"])
    p.add_argument("--init-method", type=str, default=None)
    p.add_argument("--guidance-scale", type=float, default=None)
    p.add_argument("--device", type=str, default="cuda", help="extension: 'cpu' runs the dummy model under gloo")
    return p


def main(argv=None) -> dict | None:
    args = build_parser().parse_args(argv)
    setup_logging(args.log_level)
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

    def one_video(seed: int) -> None:
        torch.manual_seed(seed)
        latent = torch.randn(shape, device=device, dtype=dtype) * noise_sigma
        for step in range(args.total_steps):
            latent = model(latent, step)

    per_rank = -(-args.num_samples // world)
    sync()
    if world > 1:
        dist.barrier()
    with torch.no_grad():
        for wi in range(args.warmup_samples):
            one_video(args.seed + rank * 10000 + wi)
    sync()
    if world > 1:
        dist.barrier()
    first, last = rank * per_rank, min((rank + 1) * per_rank, args.num_samples)
    times = []
    t0 = time.perf_counter()
    with torch.no_grad():
        for idx in range(first, last):
            ts = time.perf_counter()
            one_video(args.seed + idx)
            sync()
            times.append(time.perf_counter() - ts)
    elapsed = time.perf_counter() - t0
    mine = (elapsed, max(last - first, 0))
    everyone = [mine]
    if world > 1:
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
    results = None
    if rank == 0:
        wall = max(e for e, _ in everyone)
        measured = sum(c for _, c in everyone)
        results = {
            "mode": "data_parallel", "world_size": world, "total_steps": args.total_steps,
            "steps_per_gpu": args.total_steps, "model": args.model, "num_samples_measured": measured,
            "warmup_samples": args.warmup_samples, "samples_per_rank": per_rank, "latent_shape": list(shape),
            "first_sample_time_s": round(times[0], 4) if times else 0.0,
            "avg_sample_time_s": round(sum(times) / len(times), 4) if times else 0.0,
            "throughput_samples_per_s": round(measured / wall, 4) if wall > 0 else 0.0,
            "wall_clock_s": round(wall, 4), "per_sample_times_ms": [round(t * 1000, 2) for t in times],
        }
        LOGGER.info("BENCHMARK RESULTS (Data parallel): GPUs %d | model %s | %d samples | %.4f samples/s", world,
                    args.model, measured, results["throughput_samples_per_s"])
        print(f"BENCHMARK_JSON={json.dumps(results)}", flush=True)
    finalize_distributed()
    return results


if __name__ == "__main__":
    main()
