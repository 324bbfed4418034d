"""Karras sigma table of the reference's EulerDiscreteScheduler configuration, in closed form.

The reference builds it through diffusers at ``src/models/svd_unet.py:77-102`` with sigma_min 0.002,
sigma_max 700, rho 7, v_prediction, continuous timesteps; only three things reach the hot path:
``sigmas[n+1]`` (float32, last entry 0), ``timesteps[n] = 0.25*ln(sigma)`` and
``init_noise_sigma = sqrt(sigma_0^2 + 1)``.  Host-side, float64 ramp rounded to float32 exactly as
numpy/diffusers do.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

SIGMA_MIN = 0.002
SIGMA_MAX = 700.0
RHO = 7.0


def karras_sigmas(num_steps: int, sigma_min: float = SIGMA_MIN, sigma_max: float = SIGMA_MAX,
                  rho: float = RHO) -> np.ndarray:
    """float32 [num_steps + 1]; sigma_i = (smax^(1/rho) + i/(n-1) (smin^(1/rho) - smax^(1/rho)))^rho, then 0."""
    if num_steps <= 0:
        raise ValueError("num_steps must be positive")
    ramp = np.linspace(0.0, 1.0, num_steps)
    lo, hi = sigma_min ** (1.0 / rho), sigma_max ** (1.0 / rho)
    sig = (hi + ramp * (lo - hi)) ** rho
    return np.concatenate([sig, [0.0]]).astype(np.float32)


def continuous_timesteps(sigmas: np.ndarray) -> np.ndarray:
    """float32 [n]: 0.25 * ln(sigma), each evaluated as a 0-dim float32 torch op like diffusers does
    (numpy's float32 log differs from torch's by one ulp on some entries)."""
    import torch
    sig = torch.from_numpy(np.ascontiguousarray(sigmas[:-1], dtype=np.float32))
    return np.array([float(0.25 * x.log()) for x in sig], dtype=np.float32)


def init_noise_sigma(sigmas: np.ndarray) -> float:
    s0 = np.float32(sigmas[0])
    return float(np.sqrt(s0 * s0 + np.float32(1.0), dtype=np.float32))


def step_coefficients(sigmas: np.ndarray, step: int) -> Tuple[float, float, float, float, float]:
    """Host-side scalars of one Euler v-prediction step, in the precisions torch uses at
    svd_unet.py:382,428-437: (in_div, c_v, c_x, sigma, dt) with
      in_div = fp16(sqrt(sigma^2+1))   c_v = -sigma/sqrt(sigma^2+1)   c_x = sigma^2+1   (float32 ops)
      dt     = float32(float64(sigma_next) - float64(sigma)).
    in_div is rounded to fp16 because `latent / sqrt(sigma^2+1)` at :382 divides an fp16 tensor by a
    0-dim fp32 tensor: torch's result type is fp16 and the 0-dim operand is cast to it first."""
    s = np.float32(sigmas[step])
    s_next = np.float32(sigmas[step + 1])
    c_x = np.float32(s * s + np.float32(1.0))
    root = np.sqrt(c_x, dtype=np.float32)
    c_v = np.float32(-s / root)
    dt = np.float32(float(s_next) - float(s))
    return float(np.float16(root)), float(c_v), float(c_x), float(s), float(dt)
