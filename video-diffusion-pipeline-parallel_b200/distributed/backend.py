"""Backend string selection (reference ``src/distributed/backend.py:12-31``).

Precedence: explicit argument, then the ``PIPELINE_BACKEND`` environment variable, then ``gloo`` for
the CPU simulator and ``nccl`` otherwise.  Anything outside {nccl, gloo} is a ``ValueError``.
"""
from __future__ import annotations

import os
from typing import Optional

SUPPORTED_BACKENDS = frozenset({"nccl", "gloo"})
BACKEND_ENV_VAR = "PIPELINE_BACKEND"


def resolve_backend(preferred: Optional[str] = None, *, simulator: bool = False) -> str:
    choice = (preferred or os.environ.get(BACKEND_ENV_VAR, "")).lower()
    if not choice:
        return "gloo" if simulator else "nccl"
    if choice not in SUPPORTED_BACKENDS:
        raise ValueError(f"Unsupported backend '{choice}'.")
    return choice
