"""Process-group lifecycle behind the stage runner.

Same call contract as the reference (``src/distributed/setup.py:16-47``): ``init_distributed`` takes keyword-only
``backend, rank, world_size, init_method, timeout``, is a no-op when a group already exists, defaults to a 10-minute
timeout, and ``finalize_distributed`` tears the group down if there is one.  On top of that contract, for one
process per B200:
  * an NCCL group is bound to this rank's GPU at creation (``device_id``), so the communicator is built eagerly and
    the first latent handoff does not pay for lazy initialisation;
  * a single-node rendezvous falls back to 127.0.0.1 when no address was exported (container hostnames do not
    always resolve).
"""
from __future__ import annotations

import datetime as _dt
import logging
import os
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

LOGGER = logging.getLogger(__name__)
DEFAULT_TIMEOUT = _dt.timedelta(seconds=600)


def _group_options(backend: str, rank: int, world_size: int, init_method: Optional[str],
                   timeout: Optional[_dt.timedelta]) -> Dict[str, Any]:
    opts: Dict[str, Any] = {"backend": backend, "rank": rank, "world_size": world_size,
                            "timeout": DEFAULT_TIMEOUT if timeout is None else timeout}
    if init_method:
        opts["init_method"] = init_method
    elif "MASTER_ADDR" not in os.environ:
        # loopback only when the job is known to live on one node; a multi-node launch without an address fails right
        # away with torch's own error instead of waiting out the rendezvous timeout on 127.0.0.1
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world_size))
        if local_world == world_size:
            os.environ["MASTER_ADDR"] = "127.0.0.1"
    if backend == "nccl" and torch.cuda.is_available():
        # the launcher's LOCAL_RANK, else whatever device the caller has already selected
        local = int(os.environ["LOCAL_RANK"]) if "LOCAL_RANK" in os.environ else torch.cuda.current_device()
        opts["device_id"] = torch.device("cuda", local)
    return opts


def init_distributed(*, backend: str, rank: int, world_size: int, init_method: Optional[str] = None,
                     timeout: Optional[_dt.timedelta] = None) -> None:
    """Create the default process group unless one exists already."""
    if dist.is_initialized():
        LOGGER.debug("default process group exists; init_distributed is a no-op")
        return
    opts = _group_options(backend, rank, world_size, init_method, timeout)
    LOGGER.info("process group: %s", {k: str(v) for k, v in opts.items()})
    dist.init_process_group(**opts)


def finalize_distributed() -> None:
    """Destroy the default process group if there is one."""
    if not dist.is_initialized():
        return
    dist.destroy_process_group()
