"""ORACLE (test infrastructure): restatement of diffusers ``EulerDiscreteScheduler.set_timesteps`` for
the exact constructor arguments the reference passes at ``src/models/svd_unet.py:81-94``
(scaled_linear betas 0.00085..0.012, 1000 train steps, v_prediction, "leading" spacing,
steps_offset 1, continuous timesteps, Karras sigmas with sigma_min 0.002 / sigma_max 700).

diffusers (un-vendored dependency, >=0.20.0; authors ran 0.36.0) is absent from this image; the
algorithm below follows its published ``set_timesteps`` / ``_convert_to_karras`` step by step in the
same dtypes (numpy float64 ramp -> float32 table -> torch float32 log).  Pinned by the reference's
own statements ``sigmas[0] = 700.0`` and ``init_noise_sigma = 700.0`` (EXPERIMENT_RESULTS.md:242-243)
and by the probe values recorded in SURVEY.md section 8c.
"""
from __future__ import annotations

import numpy as np
import torch


def euler_karras_tables(num_inference_steps: int, *, num_train_timesteps: int = 1000,
                        beta_start: float = 0.00085, beta_end: float = 0.012, sigma_min: float = 0.002,
                        sigma_max: float = 700.0, steps_offset: int = 1, rho: float = 7.0):
    """Returns (sigmas float32 [n+1], timesteps float32 [n], init_noise_sigma float)."""
    n = num_inference_steps
    # scaled_linear betas -> alphas_cumprod -> the "training" sigma table (only used for interpolation,
    # which the Karras conversion below then overrides because sigma_min/max are given explicitly)
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    train_sigmas = np.array(((1 - alphas_cumprod) / alphas_cumprod) ** 0.5)
    step_ratio = num_train_timesteps // n
    timesteps = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.float32) + steps_offset
    sigmas = np.interp(timesteps, np.arange(0, len(train_sigmas)), train_sigmas)
    # _convert_to_karras
    ramp = np.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
    sig = torch.from_numpy(sigmas).to(torch.float32)
    # timestep_type="continuous" + v_prediction: t = 0.25 * ln(sigma)
    ts = torch.tensor([0.25 * s.log() for s in sig[:-1]], dtype=torch.float32)
    init_noise_sigma = float((sig[0] ** 2 + 1) ** 0.5)  # svd_unet.py:102
    return sig, ts, init_noise_sigma
