"""Image -> video generation on the native kernels: CLIP + VAE encode, the step-pipelined denoising loop, temporal decode.

Command line and behaviour of reference ``scripts/generate_video_demo.py`` (flags :33-59, flow :225-470): load and
centre-crop the input image, encode it on every rank, set the conditioning (classifier-free guidance 3.0 by default),
then for sample *i*: ``manual_seed(seed + i)``; ``randn * init_noise_sigma`` on rank 0; every rank runs its slice of
``total_steps`` Euler steps as ``model(latents, step_index)`` and hands the latent on; the last rank decodes all samples
after the loop (chunks of 4 frames, as the reference's main passes) and writes one file per sample.

What differs, all of it additive:

* the three networks are ``NativeCLIPVision`` / ``NativeVAE`` / ``StableVideoUNet`` over ``NativeUNet`` - hand-written sm_100a
  kernels end to end, no fp32 up-cast of the VAE (fp32 accumulation inside the kernels);
* ``--model-id`` is a local snapshot directory in the hub layout or ``random-init[:seed]`` (no network; the hub id maps to
  ``random-init`` when no such directory exists); ``--input-image synthetic[:seed]`` draws a test card instead of reading a file;
* the ranks run ``PipelineStage.run_many`` (stream-ordered, peer-mapped handoff on GPUs) instead of blocking
  ``dist.send`` / ``dist.recv`` per sample; ``--allow-uneven`` lifts the reference's ``total_steps % world == 0`` rule
  (25 steps on 2 / 4 / 8 GPUs), ``--transport nccl`` keeps NCCL send / recv;
* ``--schedule ring`` places stage s of sample v on rank ``(v + s) % world`` (``PipelineStage.run_many_ring``): the same
  slices and the same per-sample arithmetic (final latents SHA-256-equal), no fill / drain bubble, and every rank decodes
  and writes the samples that finish on it, each printing its own ``GENERATE_JSON=`` line;
* the noise augmentation of the conditioning image draws from a generator seeded with ``--seed`` so that every rank
  holds the same conditioning by construction (the reference relies on equal default seeds);
* outputs: ``.gif`` through PIL (and ``.mp4`` when ``imageio`` is importable - it is not in this image), ``--save-frames``
  for PNG frames, and one ``GENERATE_JSON=`` line with the measured times of every phase (device-synchronised).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import logging
import os
import time
from pathlib import Path
from typing import List, Optional, Tuple

import torch

from ..frontend import decode_latents, encode_image
from ..pipeline.pipeline import LatentSpec, PipelineConfig, PipelineStage
from ._common import setup_logging

LOGGER = logging.getLogger(__name__)
HUB_ID = "stabilityai/stable-video-diffusion-img2vid-xt"


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Generate video from image using pipeline parallel")
    p.add_argument("--input-image", type=str, required=True, help="Path to input image, or synthetic[:seed]")
    p.add_argument("--output-dir", type=str, default="outputs", help="Output directory")
    p.add_argument("--total-steps", type=int, default=25, help="Number of diffusion steps")
    p.add_argument("--num-frames", type=int, default=14, help="Number of video frames")
    p.add_argument("--fps", type=int, default=7, help="Output video FPS")
    p.add_argument("--motion-bucket-id", type=int, default=127, help="Motion bucket ID (0-255)")
    p.add_argument("--noise-aug-strength", type=float, default=0.02, help="Noise augmentation strength")
    p.add_argument("--num-samples", type=int, default=4, help="Number of samples to generate")
    p.add_argument("--seed", type=int, default=42, help="Random seed")
    p.add_argument("--height", type=int, default=576, help="Output height")
    p.add_argument("--width", type=int, default=1024, help="Output width")
    p.add_argument("--guidance-scale", type=float, default=3.0, help="CFG guidance scale (1.0 disables CFG)")
    p.add_argument("--model-id", type=str, default=HUB_ID)
    p.add_argument("--log-level", type=str, default="INFO")
    # extensions
    p.add_argument("--decode-chunk-size", type=int, default=4, help="frames per VAE decode call (reference main: 4)")
    p.add_argument("--allow-uneven", action="store_true", help="first total_steps %% world stages take one step more")
    p.add_argument("--schedule", default="fixed", choices=["fixed", "ring"],
                   help="fixed: stage s on rank s (the reference); ring: stage s of sample v on rank (v + s) %% world - no "
                        "fill / drain bubble, equal stage lengths, every rank decodes and writes the samples that end on it")
    p.add_argument("--transport", default=None, choices=["nccl", "peer"], help="stage handoff (default: peer on GPUs)")
    p.add_argument("--save-frames", action="store_true", help="also write every frame as PNG")
    p.add_argument("--no-files", action="store_true", help="generate and time only, write nothing")
    return p


def discover_distributed_info() -> Tuple[int, int, int, bool]:
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    return rank, world, int(os.environ.get("LOCAL_RANK", 0)), world > 1


# ---------------------------------------------------------------------------------------------- image in
def fit_center_crop(image, height: int, width: int):
    """Scale so the target rectangle is just covered (no distortion), then crop the centre (reference :71-89)."""
    from PIL import Image
    image = image.convert("RGB")
    w0, h0 = image.size
    s = max(width / w0, height / h0)
    w1, h1 = round(w0 * s), round(h0 * s)
    if (w1, h1) != (w0, h0):
        image = image.resize((w1, h1), Image.LANCZOS)
    x0, y0 = (w1 - width) // 2, (h1 - height) // 2
    return image.crop((x0, y0, x0 + width, y0 + height))


def synthetic_image(height: int, width: int, seed: int = 0):
    """A smooth colour test card with a few seeded blobs (no image files exist on a fresh box)."""
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    img = np.stack([xx / max(width - 1, 1), yy / max(height - 1, 1), 0.5 + 0.5 * np.sin(xx / 37.0 + yy / 53.0)], axis=-1)
    for _ in range(6):
        cx, cy, r = rng.uniform(0, width), rng.uniform(0, height), rng.uniform(0.05, 0.2) * min(height, width)
        blob = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))[..., None]
        img = img * (1 - blob) + rng.uniform(0, 1, size=3).astype(np.float32) * blob
    return Image.fromarray((img.clip(0, 1) * 255).astype(np.uint8), "RGB")


def load_and_preprocess_image(image_path: str, height: int, width: int):
    if image_path.startswith("synthetic"):
        seed = int(image_path.split(":", 1)[1]) if ":" in image_path else 0
        return synthetic_image(height, width, seed)
    from PIL import Image
    return fit_center_crop(Image.open(image_path), height, width)


def load_feature_extractor(model_id: str):
    """``CLIPImageProcessor`` (host-side resize 224 bicubic / crop / normalise, transformers): the checkpoint's
    ``feature_extractor/preprocessor_config.json`` when a local snapshot has one, else the class defaults - which are the
    SVD checkpoint's values (224, bicubic, OpenAI CLIP mean / std)."""
    from transformers import CLIPImageProcessor
    sub = os.path.join(model_id, "feature_extractor")
    if os.path.isfile(os.path.join(sub, "preprocessor_config.json")):
        return CLIPImageProcessor.from_pretrained(sub)
    return CLIPImageProcessor()


# ---------------------------------------------------------------------------------------------- frames out
def frames_to_uint8(frames: torch.Tensor):
    """``[B, 3, F, H, W]`` in [-1, 1] -> ``[F, H, W, 3]`` uint8 of the first batch element (reference :198-209).  Frames on
    a GPU go through ``svdpp_frames_to_bytes`` (one pass, bit-identical); CPU tensors through the reference's expression."""
    if frames.is_cuda and frames.shape[-1] % 4 == 0:
        from .. import native
        return native.frames_to_bytes(frames[0], rgb=True, palette=False)[0].cpu().numpy()
    f = frames[0].permute(1, 2, 3, 0)
    return ((f + 1) / 2 * 255).clamp(0, 255).to(torch.uint8).cpu().numpy()


def save_gif(frames: torch.Tensor, output_path: str, fps: int) -> None:
    """GIF of the first batch element.  Frames on a GPU are quantised there (fixed 6 x 7 x 6 colour cube, ordered dither:
    ``svdpp_frames_to_bytes``) and the host only LZW-packs them - PIL's own per-frame palette search, which CPU tensors
    still get, takes about 5 s per 25-frame 576 x 1024 video."""
    from PIL import Image
    duration = max(int(round(1000 / fps)), 1)
    if frames.is_cuda and frames.shape[-1] % 4 == 0:
        from .. import native
        idx = native.frames_to_bytes(frames[0], rgb=False, palette=True)[1].cpu().numpy()
        pal = native.cube_palette()
        imgs = []
        for a in idx:
            im = Image.fromarray(a, "P")
            im.putpalette(pal)
            imgs.append(im)
        imgs[0].save(output_path, save_all=True, append_images=imgs[1:], duration=duration, loop=0, optimize=False)
    else:
        imgs = [Image.fromarray(a) for a in frames_to_uint8(frames)]
        imgs[0].save(output_path, save_all=True, append_images=imgs[1:], duration=duration, loop=0)
    LOGGER.info("GIF saved to: %s", output_path)


def save_video(frames: torch.Tensor, output_path: str, fps: int) -> bool:
    """MP4 through imageio when it is importable (the reference's writer); False otherwise."""
    try:
        import imageio
    except ImportError:
        LOGGER.info("imageio is not installed: no %s (the GIF holds the same frames)", os.path.basename(output_path))
        return False
    imageio.mimsave(output_path, frames_to_uint8(frames), fps=fps)
    LOGGER.info("Video saved to: %s", output_path)
    return True


def save_png_frames(frames: torch.Tensor, stem: str) -> None:
    from PIL import Image
    for i, a in enumerate(frames_to_uint8(frames)):
        Image.fromarray(a).save(f"{stem}_frame{i:03d}.png")


# ---------------------------------------------------------------------------------------------- main
def _sync_time(device: torch.device) -> float:
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    return time.perf_counter()


def generate(args, *, image_encoder=None, vae=None, model=None, feature_extractor=None) -> dict:
    """The whole run; the three networks can be passed in (tests, embedding in a service) or are loaded from
    ``args.model_id``.  Returns the timing record; on the last rank ``record["frames"]`` holds the decoded videos."""
    from ..models.native_clip import NativeCLIPVision
    from ..models.native_vae import NativeVAE
    from ..models.svd_unet import StableVideoUNet

    rank, world, local_rank, distributed = discover_distributed_info()
    if not torch.cuda.is_available():
        raise RuntimeError("generate_video runs on the native CUDA kernels: no GPU visible (there is no CPU path)")
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    dtype = torch.float16
    ring = distributed and args.schedule == "ring"
    last = rank == world - 1
    if distributed:
        from ..distributed.backend import resolve_backend
        from ..distributed.setup import init_distributed
        init_distributed(backend=resolve_backend(None, simulator=False), rank=rank, world_size=world)
    model_id = args.model_id
    if model_id == HUB_ID and not os.path.isdir(model_id):
        # the reference's default is the hub id; there is no network here.  Any OTHER path that does not exist is an error
        # (the loaders raise FileNotFoundError) rather than a silent switch to noise weights.
        model_id = "random-init"
        LOGGER.warning("'%s' is not a local directory (no network here): using seeded random-init weights", args.model_id)

    t0 = _sync_time(device)
    timesteps = StableVideoUNet._default_timestep_schedule(args.total_steps)
    if image_encoder is None:
        image_encoder = NativeCLIPVision.from_pretrained(model_id, subfolder="image_encoder", torch_dtype=dtype, device=device)
    if feature_extractor is None:
        feature_extractor = load_feature_extractor(model_id)
    if vae is None:
        vae = NativeVAE.from_pretrained(model_id, subfolder="vae", torch_dtype=dtype, device=device)
    if model is None:
        model = StableVideoUNet.from_pretrained(model_id=model_id, timesteps=timesteps, torch_dtype=dtype, device=device)
        model.enable_memory_optimizations()
    rec = {"rank": rank, "world_size": world, "model_id": model_id, "load_s": _sync_time(device) - t0}
    LOGGER.info("Models loaded in %.2fs", rec["load_s"])

    image = load_and_preprocess_image(args.input_image, args.height, args.width)
    h, w = args.height // 8, args.width // 8
    t0 = _sync_time(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(args.seed)          # same conditioning on every rank by construction
    image_embeddings, image_latents = encode_image(
        image=image, image_encoder=image_encoder, feature_extractor=feature_extractor, vae=vae, device=device, dtype=dtype,
        num_frames=args.num_frames, noise_aug_strength=args.noise_aug_strength, generator=gen)
    rec["encode_s"] = _sync_time(device) - t0
    del image_encoder
    if distributed and not last and not ring:
        vae = None
    guidance = args.guidance_scale if args.guidance_scale and args.guidance_scale > 1.0 else None
    model.set_conditioning(image_embeddings=image_embeddings, image_latents=image_latents, fps=args.fps,
                           motion_bucket_id=args.motion_bucket_id, noise_aug_strength=args.noise_aug_strength,
                           guidance_scale=guidance, num_frames=args.num_frames)

    if args.num_samples > 1 and hasattr(model, "use_cuda_graph"):
        model.use_cuda_graph = True     # one graph per step index, replayed for every later sample (bit-identical to eager)
    spec = LatentSpec(shape=torch.Size((1, 4, args.num_frames, h, w)), dtype=dtype, device=device)
    sigma0 = model.init_noise_sigma

    marks: List[torch.cuda.Event] = []       # one event per sample start (+ one at the end): per-sample device time

    def mark() -> None:
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()

    def supplier(idx: int) -> torch.Tensor:
        mark()
        torch.manual_seed(args.seed + idx)
        return torch.randn(spec.shape, device=device, dtype=dtype) * sigma0

    cfg = PipelineConfig(total_steps=args.total_steps, world_size=world, rank=rank,
                         timesteps=list(range(args.total_steps)), latent_spec=spec, allow_uneven=args.allow_uneven)
    stage = PipelineStage(model=model, config=cfg, transport=args.transport or ("peer" if distributed else "nccl"))
    stage.verify_peers()       # every rank agrees on the latent and on the split, or all of them raise now
    rec["transport_note"] = stage.negotiate_transport()
    rec["transport"] = stage.transport if distributed else None
    LOGGER.info("Rank %d: steps %d to %d; generating %d samples (guidance_scale=%s)", rank, stage.step_range.start,
                stage.step_range.end - 1, args.num_samples, guidance)
    t0 = _sync_time(device)
    if ring:
        done = stage.run_many_ring(args.num_samples, input_supplier=supplier)
    else:
        done = list(enumerate(stage.run_many(args.num_samples, input_supplier=supplier if rank == 0 else None) or []))
    sample_ids, outs = [i for i, _ in done], [x for _, x in done]
    if rank == 0 and not ring:
        mark()
    rec["diffusion_s"] = _sync_time(device) - t0
    if rank == 0 and not ring:      # this rank's stage time per sample (the whole denoising loop when world_size == 1); sample 0 warms up
        rec["stage_s_by_sample"] = [round(a.elapsed_time(b) / 1000.0, 4) for a, b in zip(marks[:-1], marks[1:])]
    rec["diffusion_s_per_sample"] = rec["diffusion_s"] / args.num_samples

    frames_all: List[torch.Tensor] = []
    files: List[str] = []
    if last or (ring and outs):
        out_dir = Path(args.output_dir)
        stem_in = "synthetic" if args.input_image.startswith("synthetic") else Path(args.input_image).stem
        stamp = int(time.time())
        if not args.no_files:
            out_dir.mkdir(parents=True, exist_ok=True)
        # fingerprint of every final latent: equal across world sizes (the step pipeline is bit-identical to one GPU)
        rec["samples"] = sample_ids
        rec["latent_sha256"] = [hashlib.sha256(x.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()
                                for x in outs]
        t0 = _sync_time(device)
        for latents in outs:
            frames_all.append(decode_latents(latents, vae, args.num_frames, decode_chunk_size=args.decode_chunk_size))
        rec["decode_s"] = _sync_time(device) - t0
        rec["decode_s_per_sample"] = rec["decode_s"] / max(len(outs), 1)
        rec["frames_finite"] = bool(all(torch.isfinite(f).all().item() for f in frames_all))
        t0 = time.perf_counter()
        for idx, frames in zip(sample_ids, frames_all):
            if args.no_files:
                break
            stem = str(out_dir / f"{stem_in}_{stamp}_s{idx}_seed{args.seed + idx}")
            save_gif(frames, stem + ".gif", args.fps)
            files.append(stem + ".gif")
            if save_video(frames, stem + ".mp4", args.fps):
                files.append(stem + ".mp4")
            if args.save_frames:
                save_png_frames(frames, stem)
        rec["write_s"] = time.perf_counter() - t0
        rec["files"] = files
        LOGGER.info("=== %d samples: diffusion %.2fs (%.2fs per sample), decode %.2fs (%.2fs per sample) ===", args.num_samples,
                    rec["diffusion_s"], rec["diffusion_s_per_sample"], rec["decode_s"], rec["decode_s_per_sample"])
        print("GENERATE_JSON=" + json.dumps(rec), flush=True)
    if distributed:
        from ..distributed.setup import finalize_distributed
        finalize_distributed()
    rec["frames"] = frames_all
    return rec


def main(argv: Optional[List[str]] = None) -> dict:
    args = build_parser().parse_args(argv)
    setup_logging(args.log_level)
    return generate(args)


if __name__ == "__main__":
    main()
