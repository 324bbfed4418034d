"""Simulator mode: one latent through the step pipeline with the DummyUNet, CPU/gloo by default.

Same flags, defaults, seeding and final log line ("Final latent norm: ...") as reference
``src/modes/simulator.py:35-163`` (BASELINE config 1:
``torchrun --nproc_per_node=4 -m src.modes.simulator --total-steps 28 --latent-channels 4 --latent-frames 14
--latent-height 64 --latent-width 64``).  Extension: ``--allow-uneven`` lifts the reference's
``total_steps % world_size == 0`` requirement (25 steps on 4 ranks -> 7,6,6,6).
"""
from __future__ import annotations

import argparse
import logging

import torch

from ..distributed.backend import resolve_backend
from ..distributed.setup import finalize_distributed, init_distributed
from ..models.dummy_unet import DummyUNet
from ..pipeline.pipeline import LatentSpec, run_single_latent
from ._common import env_int, parse_dtype, setup_logging

LOGGER = logging.getLogger(__name__)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Pipeline simulator mode")
    p.add_argument("--total-steps", type=int, default=28)
    p.add_argument("--rank", type=int, default=0, help="Rank fallback when env vars missing")
    p.add_argument("--world-size", type=int, default=1)
    p.add_argument("--latent-batch", type=int, default=1)
    p.add_argument("--latent-channels", type=int, default=8)
    p.add_argument("--latent-frames", type=int, default=8)
    p.add_argument("--latent-height", type=int, default=32)
    p.add_argument("--latent-width", type=int, default=32)
    p.add_argument("--dtype", type=str, default="fp32")
    p.add_argument("--device", type=str, default="cpu",
                   help="Device string understood by torch.device (cpu, cuda:0, ...)")
    p.add_argument("--backend", type=str, default="auto", choices=["auto", "gloo", "nccl"])
    p.add_argument("--init-method", type=str, default=None)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--log-level", type=str, default="INFO")
    p.add_argument("--allow-uneven", action="store_true",
                   help="extension: allow total_steps %% world_size != 0 (first stages take one more step)")
    return p


def main(argv=None) -> float | None:
    args = build_parser().parse_args(argv)
    setup_logging(args.log_level)
    rank, world = env_int("RANK", args.rank), env_int("WORLD_SIZE", args.world_size)
    backend = resolve_backend(None if args.backend == "auto" else args.backend, simulator=True)
    init_distributed(backend=backend, rank=rank, world_size=world, init_method=args.init_method)
    dtype = parse_dtype(args.dtype)
    device = torch.device(f"cuda:{env_int('LOCAL_RANK', 0)}") if args.device == "cuda" else torch.device(args.device)
    shape = torch.Size((args.latent_batch, args.latent_channels, args.latent_frames, args.latent_height,
                        args.latent_width))
    model = DummyUNet(channels=args.latent_channels).to(device)
    latent = None
    if rank == 0:
        torch.manual_seed(args.seed)
        latent = torch.randn(shape, device=device, dtype=dtype)
    LOGGER.info("Simulator start rank=%s world_size=%s steps=%s backend=%s device=%s", rank, world,
                args.total_steps, backend, device)
    norm = None
    try:
        with torch.no_grad():
            out = run_single_latent(model=model, total_steps=args.total_steps,
                                    timesteps=list(reversed(range(args.total_steps))), world_size=world, rank=rank,
                                    latent_spec=LatentSpec(shape=shape, dtype=dtype, device=device),
                                    input_latent=latent, allow_uneven=args.allow_uneven)
        if rank == world - 1 and out is not None:
            norm = out.norm().item()
            LOGGER.info("Final latent norm: %s", norm)
    finally:
        finalize_distributed()
    return norm


if __name__ == "__main__":
    main()
