"""Image -> conditioning and latents -> frames around the denoising loop.

Same functions, arguments and conventions as the reference's generation script
(``scripts/generate_video_demo.py:92-152`` ``encode_image``, ``:154-195`` ``decode_latents``):

* CLIP image embedding ``[B, 1, 1024]`` from the feature extractor's pixel values;
* VAE latents of the conditioning image: noise augmentation in PIXEL space, ``latent_dist.mode()``, NO scaling factor,
  repeated over the frames -> ``[B, 4, F, H/8, W/8]``;
* decode: ``latents / scaling_factor``, frames decoded ``decode_chunk_size`` at a time with ``num_frames = chunk`` (the
  temporal decoder mixes the frames of a chunk), result ``[B, 3, F, H, W]`` fp32.

``vae`` / ``image_encoder`` may be the native modules (``NativeVAE``, ``NativeCLIPVision``) or any module with the
diffusers / transformers call signature; the ``force_upcast`` dance of the reference (fp32 VAE because fp16 library
convolutions overflow) only happens for modules that ask for it - the native VAE accumulates in fp32 and does not.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def encode_image(image, image_encoder, feature_extractor, vae, device: torch.device, dtype: torch.dtype, num_frames: int,
                 noise_aug_strength: float, generator: Optional[torch.Generator] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``image``: a PIL image (as in the reference) or a float tensor ``[B, 3, H, W]`` in [0, 1].
    Returns ``(image_embeddings [B, 1, 1024], image_latents [B, 4, F, H/8, W/8])``."""
    if isinstance(image, torch.Tensor):
        img01 = image.to(device=device, dtype=torch.float32)
        pixel_values = feature_extractor(img01) if feature_extractor is not None else img01
    else:
        import numpy as np
        inputs = feature_extractor(images=image, return_tensors="pt")
        pixel_values = inputs.pixel_values
        img01 = torch.from_numpy(np.asarray(image.convert("RGB"), dtype=np.float32) / 255.0).permute(2, 0, 1)[None].to(device)
    pixel_values = pixel_values.to(device, dtype=dtype)
    with torch.no_grad():
        image_embeddings = image_encoder(pixel_values).image_embeds.unsqueeze(1)          # (B, 1, 1024)
    image_tensor = ((img01 - 0.5) / 0.5).to(device, dtype=dtype)                           # Normalize([0.5], [0.5])
    if noise_aug_strength > 0:
        noise = torch.randn(image_tensor.shape, generator=generator, device=image_tensor.device, dtype=image_tensor.dtype)
        image_tensor = image_tensor + noise_aug_strength * noise
    needs_upcast = getattr(vae.config, "force_upcast", False) and vae.dtype == torch.float16
    if needs_upcast:
        vae.to(dtype=torch.float32)
        image_tensor = image_tensor.to(dtype=torch.float32)
    with torch.no_grad():
        image_latents = vae.encode(image_tensor).latent_dist.mode()        # raw VAE latents: no scaling_factor
    if needs_upcast:
        vae.to(dtype=torch.float16)
    image_latents = image_latents.to(dtype=dtype)
    return image_embeddings, image_latents.unsqueeze(2).repeat(1, 1, num_frames, 1, 1)


def decode_latents(latents: torch.Tensor, vae, num_frames: int, decode_chunk_size: int = 14) -> torch.Tensor:
    """``latents`` [B, 4, F, h, w] -> frames [B, 3, F, 8h, 8w] fp32 (reference generate_video_demo.py:154-195)."""
    latents = latents.permute(0, 2, 1, 3, 4)
    batch_size = latents.shape[0]
    latents = latents.flatten(0, 1) / vae.config.scaling_factor
    needs_upcast = getattr(vae.config, "force_upcast", False) and vae.dtype == torch.float16
    if needs_upcast:
        vae.to(dtype=torch.float32)
        latents = latents.to(dtype=torch.float32)
    frames = []
    with torch.no_grad():
        for i in range(0, latents.shape[0], decode_chunk_size):
            chunk = latents[i: i + decode_chunk_size]
            frames.append(vae.decode(chunk, num_frames=chunk.shape[0]).sample)
    frames = torch.cat(frames, dim=0)
    if needs_upcast:
        vae.to(dtype=torch.float16)
    frames = frames.reshape(batch_size, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4)
    return frames.float()
