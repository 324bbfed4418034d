"""Process-group lifecycle (reference ``src/distributed/setup.py:16-47``): idempotent init with a
10-minute default timeout, and destroy-if-initialised."""
from __future__ import annotations

import logging
from datetime import timedelta
from typing import Optional

import torch.distributed as dist

LOGGER = logging.getLogger(__name__)
DEFAULT_TIMEOUT = timedelta(minutes=10)


def init_distributed(*, backend: str, rank: int, world_size: int,
                     init_method: Optional[str] = None,
                     timeout: Optional[timedelta] = None) -> None:
    if dist.is_initialized():
        LOGGER.debug("Process group already initialized.")
        return
    kwargs = dict(backend=backend, rank=rank, world_size=world_size,
                  timeout=timeout or DEFAULT_TIMEOUT)
    if init_method:
        kwargs["init_method"] = init_method
    LOGGER.info("Initializing process group backend=%s rank=%s world_size=%s",
                backend, rank, world_size)
    dist.init_process_group(**kwargs)


def finalize_distributed() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
