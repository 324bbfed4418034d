"""Minimal stand-in for the two diffusers names the reference imports (svd_unet.py:79,129), so that
the UNMODIFIED reference wrapper can run in a container without diffusers.  Test infrastructure only:
used by tests/golden/make_golden.py (in the build container, where /root/reference exists)."""
import torch

from oracle.scheduler import euler_karras_tables


class EulerDiscreteScheduler:
    def __init__(self, **cfg):
        self.cfg = cfg
        self.sigmas = None
        self.timesteps = None

    def set_timesteps(self, n):
        c = self.cfg
        self.sigmas, self.timesteps, _ = euler_karras_tables(
            n, num_train_timesteps=c.get("num_train_timesteps", 1000), beta_start=c["beta_start"],
            beta_end=c["beta_end"], sigma_min=c["sigma_min"], sigma_max=c["sigma_max"],
            steps_offset=c.get("steps_offset", 0))
