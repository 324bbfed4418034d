"""Helpers shared by the CLI modes."""
from __future__ import annotations

import logging
import os

import torch

_DTYPES = {"float32": torch.float32, "fp32": torch.float32, "float16": torch.float16, "fp16": torch.float16,
           "bfloat16": torch.bfloat16, "bf16": torch.bfloat16}


def parse_dtype(name: str) -> torch.dtype:
    try:
        return _DTYPES[name.lower()]
    except KeyError:
        raise ValueError(f"Unsupported dtype '{name}'.") from None


def env_int(name: str, default: int) -> int:
    return int(os.environ.get(name, default))


def setup_logging(level: str) -> None:
    logging.basicConfig(level=getattr(logging, level.upper()),
                        format="%(asctime)s %(levelname)s %(name)s: %(message)s")
