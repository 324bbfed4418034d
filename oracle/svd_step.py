"""ORACLE (test infrastructure): torch restatement of the reference's per-step wrapper arithmetic,
``StableVideoUNet.set_conditioning`` / ``forward`` at ``src/models/svd_unet.py:219-279,351-439``.

Pinned against the reference itself: ``tests/golden/make_golden.py`` runs the unmodified reference
wrapper (through ``oracle/diffusers_shim``) and stores its outputs; ``tests/test_oracle.py`` requires this
restatement to reproduce them bit for bit on CPU.
"""
from __future__ import annotations

from typing import Optional

import torch

from .scheduler import euler_karras_tables


class Conditioning:
    """State built by svd_unet.py:219-279."""

    def __init__(self, image_embeddings, image_latents, *, dtype, fps=6, motion_bucket_id=127,
                 noise_aug_strength=0.02, guidance_scale=None, num_frames=14):
        if image_embeddings.dim() == 2:
            image_embeddings = image_embeddings.unsqueeze(1)
        b = image_embeddings.shape[0]
        dev = image_embeddings.device
        self.added_time_ids = torch.tensor([[fps - 1, motion_bucket_id, noise_aug_strength]], dtype=dtype,
                                           device=dev).repeat(b, 1)
        self.image_embeddings = image_embeddings.to(dtype)
        self.image_latents = image_latents.to(dtype)
        self.guidance_scale = guidance_scale
        self.cfg = guidance_scale is not None and guidance_scale > 1.0
        if self.cfg:
            self.uncond_embeddings = torch.zeros_like(self.image_embeddings)
            self.uncond_image_latents = torch.zeros_like(self.image_latents)
            gs = torch.linspace(1.0, guidance_scale, num_frames)
            self.gs = gs.view(1, 1, num_frames, 1, 1).to(dev, dtype=dtype)


def dummy_conditioning(batch_size, num_frames, height, width, device, dtype, **kw) -> Conditioning:
    """svd_unet.py:281-338: embeddings first, then latents, both ``randn`` on ``device``."""
    emb = torch.randn(batch_size, 1, 1024, device=device, dtype=dtype)
    lat = torch.randn(batch_size, 4, num_frames, height, width, device=device, dtype=dtype)
    return Conditioning(emb, lat, dtype=dtype, num_frames=num_frames, **kw)


class OracleStep:
    def __init__(self, unet, num_steps: int, dtype=torch.float16):
        self.unet = unet
        self.dtype = dtype
        self.num_steps = num_steps
        self.sigmas, self.timesteps, self.init_noise_sigma = euler_karras_tables(num_steps)

    @torch.inference_mode()
    def unet_call(self, latent_scaled, image_latents, emb, ids, t):
        x = torch.cat([latent_scaled, image_latents], dim=1).permute(0, 2, 1, 3, 4).to(self.dtype)
        return self.unet(sample=x, timestep=t, encoder_hidden_states=emb, added_time_ids=ids,
                         return_dict=False)[0]

    @torch.inference_mode()
    def __call__(self, latent: torch.Tensor, step: int, cond: Conditioning) -> torch.Tensor:
        if not (0 <= step < self.num_steps):
            raise ValueError(f"Step {step} out of range [0, {self.num_steps})")
        sigmas = self.sigmas.to(latent.device)
        sigma, sigma_next = sigmas[step], sigmas[step + 1]
        t = self.timesteps[step]
        scaled = latent / ((sigma ** 2 + 1) ** 0.5)
        if cond.cfg:
            u = self.unet_call(scaled, cond.uncond_image_latents, cond.uncond_embeddings, cond.added_time_ids, t)
            c = self.unet_call(scaled, cond.image_latents, cond.image_embeddings, cond.added_time_ids, t)
            v = u + cond.gs.permute(0, 2, 1, 3, 4) * (c - u)
        else:
            v = self.unet_call(scaled, cond.image_latents, cond.image_embeddings, cond.added_time_ids, t)
        v = v.permute(0, 2, 1, 3, 4).float()
        x = latent.float()
        s = sigma.float()
        x0 = v * (-s / (s ** 2 + 1) ** 0.5) + x / (s ** 2 + 1)
        d = (x - x0) / s
        dt = float(sigma_next) - float(sigma)
        return (x + d * dt).to(self.dtype)
