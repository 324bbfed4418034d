"""Stage runner for diffusion-step pipeline parallelism.

Interface of reference ``src/pipeline/pipeline.py`` (``LatentSpec`` :25-34, ``PipelineConfig`` :37-48,
``PipelineStage`` :54-157, ``run_single_latent`` :160-185, ``run_pipeline_latents`` :188-208): every rank
holds the whole model, owns a contiguous slice of the step schedule, receives the latent from
``rank-1``, runs ``model(latent, timesteps[i])`` for its slice and sends the latent to ``rank+1``.

What differs from the reference is *how* the handoff is issued, not what is exchanged:

* the receive buffer is a persistent pair (``LatentSpec.empty`` once per slot; the reference allocates per
  sample at ``pipeline.py:76``);
* ``run_many`` issues ``isend``/``irecv`` and waits on the *stream* (``work.wait()``), never on the host:
  the host keeps enqueueing the next video's steps while the GPU finishes the exchange.  The exchange is
  deliberately not left pending across compute: an NCCL send/recv kernel that spins for its peer occupies
  SMs, and the persistent one-CTA-per-SM kernels of this path would then run a whole extra wave.  The
  1.8 MB latent costs microseconds on NVLink against >= 100 ms of compute per stage and video;
* ``PipelineConfig.allow_uneven`` opts into ``assign_steps_uneven`` (the reference raises, Q1);
* ``PipelineStage(..., transport="peer")`` (or ``PIPELINE_TRANSPORT=peer``) replaces the NCCL send/recv by the
  peer-mapped, flag-signalled handoff of ``distributed/handoff.py``: the producer's last local step writes the latent
  straight into the next stage's receive slot over NVLink and raises its flag from the same kernel; ranks are no
  longer coupled by collective-style exchanges (CUDA devices only).

Message order, tags, shapes and the values exchanged are identical, so a mixed world of reference
and new stages interoperates.
"""
from __future__ import annotations

import logging
import os
import time
from collections.abc import Sequence
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from .step_assignment import StepRange, assign_steps, assign_steps_uneven

LOGGER = logging.getLogger(__name__)


_NVTX = os.environ.get("SVDPP_NVTX", "0") not in ("", "0")


class _nvtx:
    """NVTX range around a stage's denoising steps when SVDPP_NVTX=1 (tracing; the reference only logs
    wall-clock per step, pipeline.py:94-97).  A no-op otherwise and on CPU tensors / builds without CUDA."""

    def __init__(self, name: str):
        self.name = name
        self.on = _NVTX and torch.cuda.is_available()

    def __enter__(self):
        if self.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


@dataclass(frozen=True)
class LatentSpec:
    """Shape/dtype/device every stage agrees on out of band (reference ``pipeline.py:25-34``)."""

    shape: torch.Size
    dtype: torch.dtype
    device: torch.device

    def empty(self) -> torch.Tensor:
        return torch.empty(self.shape, dtype=self.dtype, device=self.device)


@dataclass(frozen=True)
class PipelineConfig:
    total_steps: int
    world_size: int
    rank: int
    timesteps: Sequence[int]
    latent_spec: LatentSpec
    send_tag: int = 0
    allow_uneven: bool = False  # extension: reference rejects total_steps % world_size != 0

    def __post_init__(self) -> None:
        if len(self.timesteps) != self.total_steps:
            raise ValueError("len(timesteps) must equal total_steps.")


InputSupplier = Callable[[int], torch.Tensor]


class PipelineStage:
    """One rank of the step pipeline."""

    def __init__(self, model, config: PipelineConfig, logger: Optional[logging.Logger] = None,
                 transport: Optional[str] = None) -> None:
        self.model = model
        self.config = config
        self.logger = logger or LOGGER
        self.transport = (transport or os.environ.get("PIPELINE_TRANSPORT", "nccl")).lower()
        if self.transport not in ("nccl", "peer"):
            raise ValueError("transport must be 'nccl' (dist.send / dist.recv) or 'peer' (peer-mapped slots + flags)")
        self._peer = None
        split = assign_steps_uneven if config.allow_uneven else assign_steps
        self.step_range: StepRange = split(
            total_steps=config.total_steps, world_size=config.world_size, rank=config.rank
        )
        self._local_timesteps: List[int] = list(
            config.timesteps[self.step_range.start: self.step_range.end]
        )
        self._recv_slots: List[Optional[torch.Tensor]] = [None, None]
        self._recv_turn = 0
        self._pending_send = None  # (work handle, tensor kept alive)

    # ------------------------------------------------------------------ logging
    def _log(self, message: str) -> None:
        if self.logger.isEnabledFor(logging.INFO):
            self.logger.info("[rank=%s] %s", self.config.rank, message)

    # ------------------------------------------------------------------ comm
    def _next_recv_slot(self) -> torch.Tensor:
        i = self._recv_turn
        self._recv_turn ^= 1
        if self._recv_slots[i] is None:
            self._recv_slots[i] = self.config.latent_spec.empty()
        return self._recv_slots[i]

    def _post_recv(self):
        buf = self._next_recv_slot()
        work = dist.irecv(buf, src=self.config.rank - 1, tag=self.config.send_tag)
        return work, buf

    def _recv_latent(self) -> torch.Tensor:
        """Blocking receive from ``rank-1`` (reference ``pipeline.py:75-80``).  The buffer is one of two persistent
        slots, rewritten two receives later; a stage without steps would hand the slot itself on (the model returns a
        fresh tensor per step otherwise), so it gets a copy - the reference allocates per receive."""
        self._log(f"waiting for latent from rank {self.config.rank - 1}")
        work, buf = self._post_recv()
        work.wait()
        self._log("received latent")
        return buf if self.step_range.count else buf.clone()

    def _drain_send(self) -> None:
        if self._pending_send is not None:
            self._pending_send[0].wait()
            self._pending_send = None

    def _send_latent(self, latent: torch.Tensor, *, blocking: bool = True) -> None:
        """Send to ``rank+1`` (reference ``pipeline.py:82-84``)."""
        self._log(f"sending latent to rank {self.config.rank + 1}")
        self._drain_send()
        if blocking:
            dist.send(latent, dst=self.config.rank + 1, tag=self.config.send_tag)
        else:
            work = dist.isend(latent, dst=self.config.rank + 1, tag=self.config.send_tag)
            self._pending_send = (work, latent)

    # ------------------------------------------------------------------ peer-mapped transport
    def _peer_handoff(self):
        if self._peer is None:
            from ..distributed.handoff import PeerHandoff
            spec = self.config.latent_spec
            self._peer = PeerHandoff(spec.shape, spec.dtype, spec.device)
        return self._peer

    def verify_peers(self) -> None:
        """Collective sanity check before the first handoff (failure detection): every rank publishes its
        latent shape / dtype, ``total_steps`` and step range; a mismatch raises on EVERY rank with the offending ranks named.
        The reference has no such check - ranks that disagree on the ``LatentSpec`` they agreed on "out of band"
        (pipeline.py:25-34) simply hang in ``recv`` until the 10-minute process-group timeout.  Optional, one tiny
        ``all_gather``; not called by ``run`` / ``run_many`` so that message order stays the reference's."""
        cfg = self.config
        if cfg.world_size <= 1:
            return
        shape = list(cfg.latent_spec.shape)
        if len(shape) > 8:
            raise ValueError("verify_peers supports latents of up to 8 dimensions")
        dtypes = [torch.float16, torch.float32, torch.bfloat16, torch.float64]
        code = dtypes.index(cfg.latent_spec.dtype) if cfg.latent_spec.dtype in dtypes else -1
        mine = torch.tensor([len(shape)] + shape + [0] * (8 - len(shape)) +
                            [code, cfg.total_steps, cfg.world_size, self.step_range.start, self.step_range.end],
                            dtype=torch.int64, device=cfg.latent_spec.device)
        everyone = [torch.empty_like(mine) for _ in range(cfg.world_size)]
        dist.all_gather(everyone, mine)
        rows = [t.tolist() for t in everyone]
        problems = []
        for r, row in enumerate(rows):
            if row[:12] != rows[0][:12]:
                problems.append(f"rank {r} has latent/steps/world {row[:12]} but rank 0 has {rows[0][:12]}")
        expect = 0
        for r, row in enumerate(rows):
            if row[12] != expect or row[13] < row[12]:
                problems.append(f"rank {r} owns steps [{row[12]}, {row[13]}) but the slices up to it end at {expect}")
            expect = row[13]
        if expect != rows[0][10]:
            problems.append(f"the step slices end at {expect}, not at total_steps = {rows[0][10]}")
        if problems:
            raise RuntimeError("pipeline stages disagree: " + "; ".join(problems))

    def negotiate_transport(self) -> Optional[str]:
        """Collective: set up the peer-mapped slots now and, if the symmetric-memory rendezvous fails on ANY rank, put
        every rank on NCCL send / recv instead.  Returns a note when it fell back, else None.  Optional - without it the
        first handoff sets the slots up and raises on failure."""
        if self.transport != "peer" or self.config.world_size <= 1:
            return None
        note = None
        try:
            self._peer_handoff()
            ok = 1
        except Exception as e:  # noqa: BLE001 - any failure means "not available here"
            ok, note = 0, f"peer handoff unavailable on rank {self.config.rank}: {type(e).__name__}: {e}"
        flag = torch.tensor([ok], device=self.config.latent_spec.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self.transport, self._peer = "nccl", None
            note = note or "peer handoff unavailable on another rank"
            self.logger.warning("%s; using NCCL send/recv", note)
            return note
        return None

    def _run_steps_and_hand_over(self, latent: torch.Tensor, steps: Sequence[int]) -> None:
        """The slice ``steps`` on ``latent``, its LAST step writing straight into the next rank's receive slot and
        raising that rank's flag (models with ``supports_peer_out``); other models run normally and the result is
        copied over.  Replaces ``_run_steps`` + ``_send_latent`` when ``transport == "peer"``."""
        peer = self._peer_handoff()
        steps = list(steps)
        if steps and getattr(self.model, "supports_peer_out", False):
            if len(steps) > 1:
                latent = self._run_steps(latent, steps[:-1])
            slot, handoff = peer.begin_send()
            self.model(latent, steps[-1], out=slot, handoff=handoff)
            return
        if steps:
            latent = self._run_steps(latent, steps)
        peer.send_copy(latent)

    # ------------------------------------------------------------------ compute
    def _run_local_steps(self, latent: torch.Tensor) -> torch.Tensor:
        """The inner hot loop (reference ``pipeline.py:86-98``): ``model(latent, timesteps[i])``."""
        if len(self._local_timesteps) != self.step_range.count:
            raise RuntimeError("Local timestep slice length mismatch with step range.")
        return self._run_steps(latent, self._local_timesteps)

    def _run_steps(self, latent: torch.Tensor, steps: Sequence[int]) -> torch.Tensor:
        """``model(latent, step)`` for every step of a slice; a model that offers ``forward_steps`` and asks for it
        (``use_stage_graph``) gets the whole slice in one call (one CUDA graph per stage: SURVEY 8(f) rank 2)."""
        if getattr(self.model, "use_stage_graph", False) and hasattr(self.model, "forward_steps"):
            return self.model.forward_steps(latent, steps)
        verbose = self.logger.isEnabledFor(logging.INFO)
        for step in steps:
            t0 = time.time() if verbose else 0.0
            with _nvtx(f"stage{self.config.rank}/step{step}"):
                latent = self.model(latent, step)
            if verbose:
                self._log(f"step {step} completed in {(time.time() - t0) * 1000.0:.2f} ms")
        return latent

    # ------------------------------------------------------------------ public API
    def run(self, input_latent: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """One latent through this stage; returns it on the last rank, else ``None``."""
        return self._process_single_latent(input_latent, sample_idx=None)

    def run_many(self, num_samples: int, *, input_supplier: Optional[InputSupplier] = None
                 ) -> Optional[List[torch.Tensor]]:
        if num_samples <= 0:
            raise ValueError("num_samples must be positive for pipeline execution")
        cfg = self.config
        if cfg.rank == 0 and input_supplier is None:
            raise ValueError("rank 0 requires an input_supplier when processing multiple samples")
        last = cfg.rank == cfg.world_size - 1
        outputs: List[torch.Tensor] = []
        for sample_idx in range(num_samples):
            if cfg.rank == 0:
                latent = input_supplier(sample_idx)
                if latent is None:
                    raise ValueError("rank 0 requires an input latent tensor")
                latent = latent.to(cfg.latent_spec.device)
            elif self.transport == "peer":
                latent = self._peer_handoff().recv()
            else:
                work, buf = self._post_recv()
                work.wait()
                # the slot is rewritten two samples later; the model returns a fresh tensor per step,
                # but an empty stage would forward the slot itself, hence the clone there
                latent = buf if self.step_range.count else buf.clone()
            if self.transport == "peer" and not last:
                self._run_steps_and_hand_over(latent, self._local_timesteps)
                continue
            latent = self._run_local_steps(latent)
            if last:
                outputs.append(latent)
            else:
                self._send_latent(latent, blocking=False)
                self._drain_send()      # stream-ordered wait: later kernels queue behind the send
        self._drain_send()
        return outputs if outputs else None

    def run_many_ring(self, num_samples: int, *, input_supplier: InputSupplier
                      ) -> List[tuple]:
        """Rotating stage placement (extension; same step slices, same per-video arithmetic).

        The reference pins stage s to rank s, so a stream of videos pays a fill/drain bubble of
        ``world-1`` stage times and, with an uneven split (25 steps on 8 stages), every stage waits for
        the longest one.  Every rank holds the whole model, so here stage s of video v runs on rank
        ``(v + s) % world``: in time slot t *all* ranks run stage t (equal length by construction) on
        ``world`` different videos, then all latents hop one rank along the ring.  No bubble, no
        imbalance; each video still flows through the stages in order with one handoff per boundary.

        Every rank calls ``input_supplier(v)`` for the videos it starts (v % world == rank).  Returns
        the ``(video index, final latent)`` pairs that finished on this rank.
        """
        if num_samples <= 0:
            raise ValueError("num_samples must be positive for pipeline execution")
        if input_supplier is None:
            raise ValueError("ring placement needs an input_supplier on every rank")
        cfg = self.config
        W, r = cfg.world_size, cfg.rank
        split = assign_steps_uneven if cfg.allow_uneven else assign_steps
        stages = [split(cfg.total_steps, W, s) for s in range(W)]
        outputs: List[tuple] = []
        n_batches = (num_samples + W - 1) // W
        for b in range(n_batches):
            cur: Optional[torch.Tensor] = None
            for t in range(W):
                v = b * W + (r - t) % W
                exists = v < num_samples
                if t == 0 and exists:
                    cur = input_supplier(v)
                    if cur is None:
                        raise ValueError("input_supplier returned None")
                    cur = cur.to(cfg.latent_spec.device)
                if self.transport == "peer" and t < W - 1:
                    # no collective-style exchange: my video's last step of this stage lands in the next rank's slot and
                    # raises its flag; I wait for nothing but the one latent I need next
                    if exists:
                        self._run_steps_and_hand_over(cur, cfg.timesteps[stages[t].start: stages[t].end])
                    v_in = b * W + (r - 1 - t) % W
                    cur = self._peer_handoff().recv() if v_in < num_samples else None
                    continue
                if exists:
                    cur = self._run_steps(cur, cfg.timesteps[stages[t].start: stages[t].end])
                if t == W - 1:
                    if exists:
                        # an empty last stage would append the receive slot itself (rewritten two hops later)
                        outputs.append((v, cur if stages[t].count else cur.clone()))
                    break
                # hop: my video goes to rank+1, the video of rank-1 comes to me (both at stage t -> t+1)
                v_in = b * W + (r - 1 - t) % W
                ops = []
                if exists:
                    ops.append(dist.P2POp(dist.isend, cur.contiguous(), (r + 1) % W))
                buf = None
                if v_in < num_samples:
                    buf = self._next_recv_slot()
                    ops.append(dist.P2POp(dist.irecv, buf, (r - 1) % W))
                if ops:
                    for work in dist.batch_isend_irecv(ops):
                        work.wait()
                cur = buf
        return outputs

    def _process_single_latent(self, input_latent: Optional[torch.Tensor],
                               sample_idx: Optional[int]) -> Optional[torch.Tensor]:
        """recv -> local steps -> send for one sample (reference ``pipeline.py:134-157``; called
        directly by the reference's benchmark mode, so its name and signature are kept)."""
        cfg = self.config
        prefix = f"sample {sample_idx} " if sample_idx is not None else ""
        if cfg.rank == 0:
            if input_latent is None:
                raise ValueError("rank 0 requires an input latent tensor")
            latent = input_latent.to(cfg.latent_spec.device)
            self._log(f"{prefix}input prepared")
        else:
            if input_latent is not None:
                raise ValueError("non-zero ranks should not receive an eager latent")
            latent = self._peer_handoff().recv() if self.transport == "peer" else self._recv_latent()
            self._log(f"{prefix}received latent")

        if self.transport == "peer" and cfg.rank != cfg.world_size - 1:
            self._run_steps_and_hand_over(latent, self._local_timesteps)
            return None
        latent = self._run_local_steps(latent)

        if cfg.rank == cfg.world_size - 1:
            self._log(f"{prefix}final rank completed")
            return latent
        self._send_latent(latent, blocking=True)
        return None


def run_single_latent(model, *, total_steps: int, timesteps: Sequence[int], world_size: int,
                      rank: int, latent_spec: LatentSpec, input_latent: Optional[torch.Tensor],
                      logger: Optional[logging.Logger] = None, allow_uneven: bool = False
                      ) -> Optional[torch.Tensor]:
    """All ranks call this once per latent; rank 0 passes the input (reference ``pipeline.py:160``)."""
    config = PipelineConfig(total_steps=total_steps, world_size=world_size, rank=rank,
                            timesteps=timesteps, latent_spec=latent_spec, allow_uneven=allow_uneven)
    return PipelineStage(model=model, config=config, logger=logger).run(input_latent=input_latent)


def run_pipeline_latents(model, *, total_steps: int, timesteps: Sequence[int], world_size: int,
                         rank: int, latent_spec: LatentSpec, num_samples: int,
                         input_supplier: Optional[InputSupplier],
                         logger: Optional[logging.Logger] = None, allow_uneven: bool = False
                         ) -> Optional[List[torch.Tensor]]:
    """Stream ``num_samples`` latents through the pipeline (reference ``pipeline.py:188-208``)."""
    config = PipelineConfig(total_steps=total_steps, world_size=world_size, rank=rank,
                            timesteps=timesteps, latent_spec=latent_spec, allow_uneven=allow_uneven)
    return PipelineStage(model=model, config=config, logger=logger).run_many(
        num_samples, input_supplier=input_supplier)
