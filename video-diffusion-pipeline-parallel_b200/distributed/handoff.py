"""Stage-to-stage latent handoff over peer-mapped memory (NVLink 5 / NVSwitch), the B200 replacement for the blocking
``dist.send`` / ``dist.recv`` pair of reference ``src/pipeline/pipeline.py:75-84``.

What is exchanged, in which order and between which ranks is unchanged (one latent per video per stage boundary, from
the rank that ran stage s to the rank that runs stage s+1); only the transport differs:

* every rank owns two receive slots and four flags in SYMMETRIC memory (``torch.distributed._symmetric_memory``: the
  same allocation mapped into every peer's address space);
* the producer's last local step writes its result straight into the consumer's slot - the Euler kernel's ``out`` is
  the peer-mapped pointer, so the 1.8 MB latent crosses NVLink as that kernel's ordinary stores - and the same kernel
  raises the consumer's ``ready`` flag when its stores are visible (``svdpp_euler_vpred_step_signal``);
* the consumer's stream waits for the flag (``svdpp_flag_wait``: one spinning thread, bounded by a timeout), takes a
  private copy of the slot and immediately hands the slot back by raising the producer's ``ack`` flag
  (``svdpp_flag_set``), which the producer checks before it reuses that slot two videos later.

There is no NCCL kernel on the data path, no host synchronisation and no lock-step: a rank never waits for anything but
the one latent it needs next.  Everything is stream-ordered and CUDA-graph capturable (the flag values are 0 / 1).

Models that accept ``forward(latent, step, out=..., handoff=...)`` (``supports_peer_out``: StableVideoUNet) get the fused
path; any other model runs its last step normally and the result is copied into the peer slot by a device copy followed
by ``svdpp_flag_set``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

READY0, ACK0, N_FLAGS = 0, 2, 16       # flag indices inside each rank's flag block (uint32 each)


class PeerHandoff:
    """Symmetric receive slots + flags of one rank, and the peer-mapped views of its ring neighbours'."""

    def __init__(self, shape: Sequence[int], dtype: torch.dtype, device: torch.device, group=None, timeout_s: int = 600):
        import torch.distributed._symmetric_memory as symm
        from .. import native
        if device.type != "cuda":
            raise native.NativeError("peer-mapped handoff needs CUDA devices (use the default send/recv transport on CPU)")
        native.load()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        self.timeout_s = timeout_s
        self.nxt, self.prv = (self.rank + 1) % self.world, (self.rank - 1) % self.world
        self._slots = symm.empty((2,) + self.shape, dtype=dtype, device=device)
        self._flags = symm.empty((N_FLAGS,), dtype=torch.int32, device=device)
        self._flags.zero_()
        self._flags[ACK0:ACK0 + 2] = 1                  # both of the next rank's slots start out free
        torch.cuda.synchronize(device)
        self._h_slots = symm.rendezvous(self._slots, self.group)
        self._h_flags = symm.rendezvous(self._flags, self.group)
        dist.barrier(self.group)
        self.peer_slots = self._h_slots.get_buffer(self.nxt, (2,) + self.shape, dtype)      # next rank's receive slots
        self._nxt_flags = self._h_flags.get_buffer(self.nxt, (N_FLAGS,), torch.int32)       # next rank's flag block
        self._prv_flags = self._h_flags.get_buffer(self.prv, (N_FLAGS,), torch.int32)       # previous rank's flag block
        self._done = torch.zeros(2, dtype=torch.int32, device=device)                       # local completion counters
        self._send_turn = 0
        self._recv_turn = 0
        self._sent = [0, 0]

    # ------------------------------------------------------------------ producer side
    def begin_send(self):
        """Claim the next rank's next receive slot: waits (on the stream) until that rank has handed it back.  Returns
        ``(slot tensor on the peer, handoff tuple)`` for ``model(latent, step, out=slot, handoff=handoff)``."""
        from .. import native
        k = self._send_turn
        self._send_turn ^= 1
        # ack[k] lives in MY flag block; the next rank raises it once it has copied the slot's previous content
        native.flag_wait(self._flags[ACK0 + k:].data_ptr(), 1, reset_to=0, timeout_s=self.timeout_s)
        handoff = (self._done[k:].data_ptr(), self._nxt_flags[READY0 + k:].data_ptr(), 1)
        return self.peer_slots[k], handoff

    def send_copy(self, latent: torch.Tensor) -> None:
        """Un-fused send for models without ``out=``: device copy into the peer slot, then the flag."""
        from .. import native
        slot, handoff = self.begin_send()
        slot.copy_(latent)
        native.flag_set(handoff[1], 1)

    # ------------------------------------------------------------------ consumer side
    def recv(self) -> torch.Tensor:
        """Stream-ordered receive: waits for the previous rank's flag, copies the slot and hands it back."""
        from .. import native
        k = self._recv_turn
        self._recv_turn ^= 1
        native.flag_wait(self._flags[READY0 + k:].data_ptr(), 1, reset_to=0, timeout_s=self.timeout_s)
        latent = self._slots[k].clone()
        native.flag_set(self._prv_flags[ACK0 + k:].data_ptr(), 1)
        return latent

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
