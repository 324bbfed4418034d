"""StableVideoUNet: ``forward(latent, step) -> latent`` adapter around the SVD UNet operator.

Same public surface as reference ``src/models/svd_unet.py`` (constructor :42-75, ``from_pretrained``
:104-164, ``enable_memory_optimizations`` :166-194, ``init_noise_sigma`` :196-199,
``_default_timestep_schedule`` :201-217, ``set_conditioning`` :219-279, ``set_dummy_conditioning``
:281-338, ``clear_conditioning`` :340-349, ``forward`` :351-439) and the same error behaviour
(``RuntimeError`` without conditioning, ``ValueError`` for a step out of range).

What runs underneath differs: one denoising step is
  pack (scale_model_input + cat + permute, one kernel) -> UNet -> CFG combine + Euler update (one kernel)
on libsvdpp.so kernels.  With a ``NativeUNet`` the activations never leave channels-last layout and the
two classifier-free-guidance branches run as one batch of 2 (the reference runs them sequentially,
:384-411); with any other UNet module (boundary B2) the operator is called exactly as the reference
does.  Sigma arithmetic is done on the host once (the reference syncs twice per step at :436).
Optionally the whole step is captured in a CUDA graph per step index (``use_cuda_graph``).

There is no CPU path: ``forward`` raises ``NativeError`` for non-CUDA latents or a missing library.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import scheduler as sched


class StableVideoUNet(nn.Module):
    def __init__(self, unet: nn.Module, timesteps: Sequence[int], dtype: torch.dtype = torch.float16,
                 num_train_timesteps: int = 1000) -> None:
        super().__init__()
        self.unet = unet
        self.timesteps = list(timesteps)
        self.dtype = dtype
        self.num_train_timesteps = num_train_timesteps
        self._init_scheduler()
        self.register_buffer("_image_embeddings", None, persistent=False)
        self.register_buffer("_added_time_ids", None, persistent=False)
        self.register_buffer("_image_latents", None, persistent=False)
        self._conditioning_set = False
        self._guidance_scale = None
        self._uncond_embeddings = None
        self._uncond_image_latents = None
        self._guidance_scale_tensor = None
        # execution options (extensions)
        self.use_cuda_graph = False
        self.use_stage_graph = False          # forward_steps(): one CUDA graph for a whole stage's step slice
        self._graphs: Dict[tuple, tuple] = {}
        self._graph_pool = None
        self._warm: set = set()
        self._cfg_cache = None

    # ------------------------------------------------------------------ scheduler
    def _init_scheduler(self) -> None:
        """Karras table of the reference's EulerDiscreteScheduler config (svd_unet.py:77-102)."""
        sig = sched.karras_sigmas(len(self.timesteps))
        self._sigmas_np = sig
        self.register_buffer("sigmas", torch.from_numpy(sig.copy()), persistent=False)
        self.scheduler_timesteps = torch.from_numpy(sched.continuous_timesteps(sig))  # stays on CPU, as :99
        self._init_noise_sigma = sched.init_noise_sigma(sig)

    @property
    def init_noise_sigma(self) -> float:
        return self._init_noise_sigma

    @staticmethod
    def _default_timestep_schedule(num_steps: int, num_train_timesteps: int = 1000) -> list:
        ratio = num_train_timesteps // num_steps
        return list(range(num_train_timesteps - 1, -1, -ratio))[:num_steps]

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_pretrained(cls, model_id: str = "stabilityai/stable-video-diffusion-img2vid-xt",
                        timesteps: Optional[Sequence[int]] = None, torch_dtype: torch.dtype = torch.float16,
                        enable_memory_efficient_attention: bool = True, enable_sliced_attention: bool = False,
                        attention_slice_size="auto", **kwargs) -> "StableVideoUNet":
        """Build a NativeUNet-backed wrapper.

        ``model_id`` is either a local directory holding ``unet/diffusion_pytorch_model*.safetensors``
        (diffusers layout), or ``"random-init"`` / ``"random-init:<seed>"`` for seeded default-initialised
        weights of the SVD-XT architecture (no hub access exists here).  The attention/slicing flags
        are accepted for signature compatibility; the native attention is always fused."""
        import os

        from .native_unet import NativeUNet
        from .svd_weights import random_state_dict

        device = kwargs.pop("device", "cuda")
        config = kwargs.pop("config", None)
        orchestrator = kwargs.pop("orchestrator", None)   # "c" (default: csrc/unet.cu) or "python" (per-kernel ctypes launches)
        if model_id.startswith("random-init"):
            seed = int(model_id.split(":", 1)[1]) if ":" in model_id else 0
            sd = random_state_dict(config, seed=seed, device=device)
        elif os.path.isdir(model_id):
            from safetensors.torch import load_file
            unet_dir = os.path.join(model_id, "unet") if os.path.isdir(os.path.join(model_id, "unet")) else model_id
            files = sorted(f for f in os.listdir(unet_dir) if f.endswith(".safetensors"))
            if not files:
                raise FileNotFoundError(f"no .safetensors file under {unet_dir}")
            pick = [f for f in files if "fp16" in f] or files
            sd = load_file(os.path.join(unet_dir, pick[0]), device=str(device))
        else:
            raise FileNotFoundError(
                f"'{model_id}' is not a local checkpoint directory and there is no network access; "
                "use a local path or 'random-init[:seed]'")
        unet = NativeUNet(sd, config=config, device=device, orchestrator=orchestrator)
        del sd
        if timesteps is None:
            timesteps = cls._default_timestep_schedule(num_steps=25)
        return cls(unet=unet, timesteps=timesteps, dtype=torch_dtype).to(device)

    def enable_memory_optimizations(self) -> None:
        """Reference tries xformers / flash / checkpointing toggles (svd_unet.py:166-194); the native
        operator needs none of them, foreign operators get the same best-effort calls."""
        for name, args in (("enable_xformers_memory_efficient_attention", ()),
                           ("set_attention_backend", ("flash_attention_2",))):
            fn = getattr(self.unet, name, None)
            if fn is not None:
                try:
                    fn(*args)
                    return
                except Exception:  # noqa: BLE001 - best effort, as in the reference
                    pass

    # ------------------------------------------------------------------ conditioning
    def set_conditioning(self, image_embeddings: torch.Tensor, image_latents: torch.Tensor, fps: int = 6,
                         motion_bucket_id: int = 127, noise_aug_strength: float = 0.02,
                         guidance_scale: Optional[float] = None, num_frames: int = 14) -> None:
        if image_embeddings.dim() == 2:
            image_embeddings = image_embeddings.unsqueeze(1)
        batch = image_embeddings.shape[0]
        device = image_embeddings.device
        self._added_time_ids = torch.tensor([[fps - 1, motion_bucket_id, noise_aug_strength]], dtype=self.dtype,
                                            device=device).repeat(batch, 1)
        self._image_embeddings = image_embeddings.to(self.dtype)
        self._image_latents = image_latents.to(self.dtype).contiguous()
        self._conditioning_set = True
        self._guidance_scale = guidance_scale
        if guidance_scale is not None and guidance_scale > 1.0:
            self._uncond_embeddings = torch.zeros_like(self._image_embeddings)
            self._uncond_image_latents = torch.zeros_like(self._image_latents)
            gs = torch.linspace(1.0, guidance_scale, num_frames)
            self._guidance_scale_tensor = gs.view(1, 1, num_frames, 1, 1).to(device, dtype=self.dtype)
        else:
            self._uncond_embeddings = None
            self._uncond_image_latents = None
            self._guidance_scale_tensor = None
        self._cfg_cache = None
        self._graphs.clear()
        self._warm.clear()

    def set_dummy_conditioning(self, batch_size: int, num_frames: int, height: int, width: int,
                               device: torch.device, fps: int = 6, motion_bucket_id: int = 127,
                               noise_aug_strength: float = 0.02, guidance_scale: Optional[float] = None) -> None:
        # same draw order as the reference (:310-328): embeddings, then latents
        emb = torch.randn(batch_size, 1, 1024, device=device, dtype=self.dtype)
        lat = torch.randn(batch_size, 4, num_frames, height, width, device=device, dtype=self.dtype)
        self.set_conditioning(emb, lat, fps=fps, motion_bucket_id=motion_bucket_id,
                              noise_aug_strength=noise_aug_strength, guidance_scale=guidance_scale,
                              num_frames=num_frames)

    def clear_conditioning(self) -> None:
        self._image_embeddings = None
        self._added_time_ids = None
        self._image_latents = None
        self._conditioning_set = False
        self._guidance_scale = None
        self._uncond_embeddings = None
        self._uncond_image_latents = None
        self._guidance_scale_tensor = None
        self._cfg_cache = None
        self._graphs.clear()
        self._warm.clear()

    @property
    def _cfg_on(self) -> bool:
        return self._guidance_scale is not None and self._guidance_scale > 1.0

    def _check_conditioning_shapes(self, latent: torch.Tensor) -> None:
        """The kernels index the conditioning tensors with the LATENT's shape (raw pointers, no length arguments), so a
        mismatch would read out of bounds where the reference gets a torch broadcast / cat error (svd_unet.py:387,410)."""
        B, C, F, H, W = latent.shape
        if tuple(self._image_latents.shape) != (B, C, F, H, W):
            raise ValueError(f"image_latents {tuple(self._image_latents.shape)} do not match the latent {tuple(latent.shape)}")
        emb = self._image_embeddings
        if emb.dim() != 3 or emb.shape[0] != B or emb.shape[1] != 1:
            raise ValueError(f"image_embeddings must be [B={B}, 1, D] (one CLIP token per sample), got {tuple(emb.shape)}")
        if self._added_time_ids.shape[0] != B:
            raise ValueError(f"added_time_ids batch {self._added_time_ids.shape[0]} != latent batch {B}")
        if self._cfg_on and self._guidance_scale_tensor.numel() != F:
            raise ValueError(f"guidance ramp has {self._guidance_scale_tensor.numel()} frames (set_conditioning num_frames) "
                             f"but the latent has {F}")

    # ------------------------------------------------------------------ one step
    def _step_native(self, latent: torch.Tensor, step: int, out: Optional[torch.Tensor] = None, handoff=None
                     ) -> torch.Tensor:
        """pack -> NativeUNet (channels-last, CFG batched) -> CFG + Euler.  ``out`` (optional) receives the result - it
        may be the next stage's peer-mapped receive slot, in which case ``handoff`` makes the Euler kernel raise that
        stage's flag (distributed/handoff.py)."""
        from .. import native
        B, C, F, H, W = latent.shape
        in_div, c_v, c_x, sigma, dt = sched.step_coefficients(self._sigmas_np, step)
        dev = latent.device
        M = B * F * H * W
        strides = (C * F * H * W, H * W, F * H * W)  # (b, f, c) element strides of [B,C,F,H,W]
        cfg = self._cfg_on
        nb = 2 * B if cfg else B
        if cfg and self._cfg_cache is None:
            self._cfg_cache = (torch.cat([self._uncond_embeddings, self._image_embeddings]).contiguous(),
                               torch.cat([self._added_time_ids, self._added_time_ids]).contiguous(),
                               self._guidance_scale_tensor.reshape(-1).contiguous())
        if hasattr(self.unet, "step_native") and getattr(self.unet, "orchestrator", "") == "c":
            # the whole step behind the C ABI: one svdpp_unet_step call (pack, UNet, guidance + Euler)
            enc, ids, gs = self._cfg_cache if cfg else (self._image_embeddings, self._added_time_ids, None)
            return self.unet.step_native(torch.empty_like(latent) if out is None else out, latent, self._image_latents,
                                         self._uncond_image_latents if cfg else None, enc, ids, gs,
                                         timestep=float(self.scheduler_timesteps[step]), in_div=in_div, c_v=c_v, c_x=c_x,
                                         sigma=sigma, dt=dt, handoff=handoff)
        x_in = torch.empty((nb * F * H * W, 2 * C), dtype=torch.float16, device=dev)
        if cfg:
            native.pack_unet_input(x_in[:M], latent, strides, C, in_div, self._uncond_image_latents, strides, C,
                                   B=B, F=F, H=H, W=W)
            native.pack_unet_input(x_in[M:], latent, strides, C, in_div, self._image_latents, strides, C,
                                   B=B, F=F, H=H, W=W)
            enc, ids, gs = self._cfg_cache
        else:
            native.pack_unet_input(x_in, latent, strides, C, in_div, self._image_latents, strides, C,
                                   B=B, F=F, H=H, W=W)
            enc, ids, gs = self._image_embeddings, self._added_time_ids, None
        t_dev = torch.full((nb,), float(self.scheduler_timesteps[step]), dtype=torch.float32, device=dev)
        v = self.unet.forward_nhwc(x_in, t_dev, enc, ids, nb, F, H, W)
        out = torch.empty_like(latent) if out is None else out
        if cfg:
            native.euler_vpred_step(out, latent, v[:M], v_cond=v[M:], gs=gs, v_nhwc=True, c_v=c_v, c_x=c_x,
                                    sigma=sigma, dt=dt, handoff=handoff)
        else:
            native.euler_vpred_step(out, latent, v, v_nhwc=True, c_v=c_v, c_x=c_x, sigma=sigma, dt=dt, handoff=handoff)
        return out

    def _step_foreign(self, latent: torch.Tensor, step: int, out: Optional[torch.Tensor] = None, handoff=None
                      ) -> torch.Tensor:
        """Any other UNet module: call the operator exactly as the reference does (B2)."""
        from .. import native
        B, C, F, H, W = latent.shape
        in_div, c_v, c_x, sigma, dt = sched.step_coefficients(self._sigmas_np, step)
        strides = (C * F * H * W, H * W, F * H * W)
        timestep = self.scheduler_timesteps[step]

        def call(image_latents, emb):
            sample = torch.empty((B, F, 2 * C, H, W), dtype=torch.float16, device=latent.device)
            native.pack_unet_input(sample, latent, strides, C, in_div, image_latents, strides, C, B=B, F=F, H=H, W=W,
                                   out_bfchw=True)
            return self.unet(sample=sample, timestep=timestep, encoder_hidden_states=emb,
                             added_time_ids=self._added_time_ids, return_dict=False)[0].contiguous()

        out = torch.empty_like(latent) if out is None else out
        if self._cfg_on:
            u = call(self._uncond_image_latents, self._uncond_embeddings)
            c = call(self._image_latents, self._image_embeddings)
            native.euler_vpred_step(out, latent, u, v_cond=c, gs=self._guidance_scale_tensor.reshape(-1).contiguous(),
                                    v_nhwc=False, c_v=c_v, c_x=c_x, sigma=sigma, dt=dt, handoff=handoff)
        else:
            v = call(self._image_latents, self._image_embeddings)
            native.euler_vpred_step(out, latent, v, v_nhwc=False, c_v=c_v, c_x=c_x, sigma=sigma, dt=dt, handoff=handoff)
        return out

    supports_peer_out = True      # forward(latent, step, out=..., handoff=...): see distributed/handoff.py

    def _step(self, latent: torch.Tensor, step: int, out: Optional[torch.Tensor] = None, handoff=None) -> torch.Tensor:
        if hasattr(self.unet, "forward_nhwc"):
            return self._step_native(latent, step, out, handoff)
        return self._step_foreign(latent, step, out, handoff)

    @torch.inference_mode()
    def forward_steps(self, latent: torch.Tensor, steps: Sequence[int]) -> torch.Tensor:
        """All of ``steps`` on one latent: ``for s in steps: latent = self(latent, s)`` - what
        ``PipelineStage._run_local_steps`` does with a stage's slice of the schedule (reference pipeline.py:86-98).
        With ``use_stage_graph`` the whole slice (every launch of every step, ~740 per step) is ONE CUDA graph per
        (slice, shape), replayed with a single launch call per video and stage."""
        steps = tuple(int(s) for s in steps)
        if not steps:
            return latent
        if not self.use_stage_graph:
            for s in steps:
                latent = self.forward(latent, s)
            return latent
        if not self._conditioning_set:
            raise RuntimeError("Conditioning not set. Call set_conditioning() or "
                               "set_dummy_conditioning() before forward().")
        for s in steps:
            if not (0 <= s < len(self.timesteps)):
                raise ValueError(f"Step {s} out of range [0, {len(self.timesteps)})")
        from ..native import NativeError
        if not latent.is_cuda:
            raise NativeError("StableVideoUNet.forward needs a CUDA latent: this build has no CPU path")
        latent = latent.to(torch.float16).contiguous()
        self._check_conditioning_shapes(latent)
        shape_key = tuple(latent.shape)
        if shape_key not in self._warm:           # first call per shape runs eagerly (fills caches, sizes workspaces)
            self._warm.add(shape_key)
            for s in steps:
                latent = self._step(latent, s)
            return latent
        key = ("stage", steps, shape_key)
        if key not in self._graphs:
            from .. import native
            g_in = latent.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            before = native.LAUNCHES
            with torch.cuda.graph(graph, pool=self._graph_pool):
                x = g_in
                for s in steps:
                    x = self._step(x, s)
            n_kernels = native.LAUNCHES - before
            native.LAUNCHES = before
            self._graphs[key] = (graph, g_in, x, n_kernels)
        graph, g_in, g_out, n_kernels = self._graphs[key]
        g_in.copy_(latent)
        graph.replay()
        from .. import native
        native.LAUNCHES += n_kernels
        return g_out.clone()

    @torch.inference_mode()
    def forward(self, latent: torch.Tensor, step: int, out: Optional[torch.Tensor] = None, handoff=None) -> torch.Tensor:
        """One denoising step with the scheduler update (reference svd_unet.py:351-439).  Extension: ``out`` = tensor
        that receives the result (e.g. the next pipeline stage's peer-mapped receive slot) and ``handoff`` =
        ``(done_counter_ptr, ready_flag_ptr, value)`` for the flag the Euler kernel raises once ``out`` is complete."""
        if not self._conditioning_set:
            raise RuntimeError("Conditioning not set. Call set_conditioning() or "
                               "set_dummy_conditioning() before forward().")
        if not (0 <= step < len(self.timesteps)):
            raise ValueError(f"Step {step} out of range [0, {len(self.timesteps)})")
        from ..native import NativeError
        if not latent.is_cuda:
            raise NativeError("StableVideoUNet.forward needs a CUDA latent: this build has no CPU path")
        latent = latent.to(torch.float16).contiguous()
        if latent.dim() != 5:
            raise ValueError(f"latent must be [B, C, F, H, W], got {tuple(latent.shape)}")
        self._check_conditioning_shapes(latent)
        if not self.use_cuda_graph:
            return self._step(latent, step, out, handoff)
        key = (step, tuple(latent.shape), None if out is None else out.data_ptr(), handoff)
        shape_key = tuple(latent.shape)
        if shape_key not in self._warm:           # first call per shape runs eagerly (fills caches)
            self._warm.add(shape_key)
            return self._step(latent, step, out, handoff)
        if key not in self._graphs:
            g_in = latent.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            from .. import native
            before = native.LAUNCHES
            with torch.cuda.graph(graph, pool=self._graph_pool):
                g_out = self._step(g_in, step, out, handoff)
            n_kernels = native.LAUNCHES - before
            native.LAUNCHES = before          # capture launches nothing; replays are counted below
            self._graphs[key] = (graph, g_in, g_out, n_kernels)
        graph, g_in, g_out, n_kernels = self._graphs[key]
        g_in.copy_(latent)
        graph.replay()
        from .. import native
        native.LAUNCHES += n_kernels
        return g_out.clone() if out is None else out
