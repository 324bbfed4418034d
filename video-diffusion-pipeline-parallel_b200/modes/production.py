"""Production mode: stream videos through the step pipeline on NCCL with the native SVD UNet.

Same command line as reference ``src/modes/production.py:20-47`` (``--total-steps`` (required), ``--latent-shape B C F H
W``, ``--timesteps``, ``--num-samples``, ``--seed``, ``--log-level``, ``--backend``, ``--init-method``, ``--model-id``,
``--fps``, ``--motion-bucket-id``, ``--enable-memory-opt``, ``--attention-slicing``) and the same behaviour
(:62-145): one process per GPU under torchrun, dummy image conditioning, sample *i* = ``manual_seed(seed + i); randn *
init_noise_sigma`` on rank 0, ``run_pipeline_latents`` over the whole stream, no result printed by the reference.
Differences, all additive: ``--model-id`` is a local diffusers-layout directory or ``random-init[:seed]`` (the hub id
maps to ``random-init``: no network), ``--allow-uneven`` / ``--schedule ring`` / ``--guidance-scale`` select the
extensions, and the last rank logs the norm of the final latent of every video (``--save-latents DIR`` writes them).
The reference's default schedule is DESCENDING values used as sigma indices; ``--ascending`` walks the
noise schedule the physically meaningful way.
"""
from __future__ import annotations

import argparse
import logging
import os
import time
from collections.abc import Sequence

import torch

from ..distributed.backend import resolve_backend
from ..distributed.setup import finalize_distributed, init_distributed
from ..models.svd_unet import StableVideoUNet
from ..pipeline.pipeline import LatentSpec, PipelineConfig, PipelineStage
from ._common import setup_logging

LOGGER = logging.getLogger(__name__)
HUB_ID = "stabilityai/stable-video-diffusion-img2vid-xt"


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Production pipeline mode")
    p.add_argument("--total-steps", type=int, required=True)
    p.add_argument("--latent-shape", type=int, nargs=5, metavar=("B", "C", "F", "H", "W"))
    p.add_argument("--timesteps", type=int, nargs="+", help="Explicit timestep schedule")
    p.add_argument("--num-samples", type=int, default=1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--log-level", type=str, default="INFO")
    p.add_argument("--backend", type=str, default="auto")
    p.add_argument("--init-method", type=str, default=None)
    p.add_argument("--model-id", type=str, default=HUB_ID)
    p.add_argument("--fps", type=int, default=6, help="Frames per second for conditioning")
    p.add_argument("--motion-bucket-id", type=int, default=127, help="Motion bucket ID (0-255)")
    p.add_argument("--enable-memory-opt", action="store_true", help="accepted; the native attention is always fused")
    p.add_argument("--attention-slicing", action="store_true", help="accepted; no effect on the native operator")
    # extensions
    p.add_argument("--guidance-scale", type=float, default=None)
    p.add_argument("--allow-uneven", action="store_true")
    p.add_argument("--schedule", default="fixed", choices=["fixed", "ring"])
    p.add_argument("--ascending", action="store_true", help="default schedule 0..T-1 instead of the reference's T-1..0")
    p.add_argument("--save-latents", type=str, default=None, help="directory for the final latents (last rank)")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    setup_logging(args.log_level)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", 0)))
    backend = resolve_backend(None if args.backend == "auto" else args.backend, simulator=False)
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    init_distributed(backend=backend, rank=rank, world_size=world, init_method=args.init_method)

    if args.timesteps:
        timesteps: Sequence[int] = args.timesteps
    elif args.ascending:
        timesteps = list(range(args.total_steps))
    else:
        timesteps = list(range(args.total_steps - 1, -1, -1))
    model_id = "random-init" if args.model_id == HUB_ID and not os.path.isdir(args.model_id) else args.model_id
    LOGGER.info("Loading model from %s on rank %d", model_id, rank)
    model = StableVideoUNet.from_pretrained(model_id=model_id, timesteps=timesteps, torch_dtype=torch.float16,
                                            enable_memory_efficient_attention=args.enable_memory_opt,
                                            enable_sliced_attention=args.attention_slicing, device=device)
    if args.enable_memory_opt:
        model.enable_memory_optimizations()
    shape = tuple(args.latent_shape) if args.latent_shape else (1, 4, 14, 64, 64)
    model.set_dummy_conditioning(batch_size=shape[0], num_frames=shape[2], height=shape[3], width=shape[4],
                                 device=device, fps=args.fps, motion_bucket_id=args.motion_bucket_id,
                                 guidance_scale=args.guidance_scale)
    model.use_cuda_graph = True
    LOGGER.info("Model loaded and conditioning set on rank %d", rank)
    spec = LatentSpec(shape=torch.Size(shape), dtype=torch.float16, device=device)
    sigma0 = model.init_noise_sigma

    def supplier(idx: int) -> torch.Tensor:
        torch.manual_seed(args.seed + idx)
        return torch.randn(spec.shape, device=device, dtype=spec.dtype) * sigma0

    cfg = PipelineConfig(total_steps=args.total_steps, world_size=world, rank=rank, timesteps=timesteps,
                         latent_spec=spec, allow_uneven=args.allow_uneven)
    stage = PipelineStage(model=model, config=cfg)
    stage.verify_peers()       # ranks that disagree on the latent or the split raise now instead of hanging in recv
    t0 = time.perf_counter()
    if args.schedule == "ring" and world > 1:
        done = stage.run_many_ring(args.num_samples, input_supplier=supplier)
    else:
        outs = stage.run_many(args.num_samples, input_supplier=supplier if rank == 0 else None)
        done = list(enumerate(outs or []))
    torch.cuda.synchronize(device)
    for idx, lat in done:
        LOGGER.info("video %d final latent norm %.4f (rank %d, %.2f s since start)", idx, lat.float().norm().item(), rank,
                    time.perf_counter() - t0)
        if args.save_latents:
            os.makedirs(args.save_latents, exist_ok=True)
            torch.save(lat.cpu(), os.path.join(args.save_latents, f"latent_{idx:05d}.pt"))
    finalize_distributed()
    return done


if __name__ == "__main__":
    main()
