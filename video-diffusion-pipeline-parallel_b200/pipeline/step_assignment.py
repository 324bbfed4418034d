"""Diffusion-step -> pipeline-stage assignment.

Mirrors reference ``src/pipeline/step_assignment.py:12-69`` (``StepRange``, ``assign_steps``) bit for
bit, including its ``ValueError`` cases, and adds the uneven split the reference rejects
(``assign_steps_uneven``) which BASELINE configs 1/3/5 need (25 steps on
2/4/8 stages).  Pure Python, no torch import, like the reference (``step_assignment.py:3-4``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List


@dataclass(frozen=True)
class StepRange:
    """Half-open interval ``[start, end)`` of schedule positions owned by one stage."""

    start: int
    end: int

    def __post_init__(self) -> None:
        if self.start < 0 or self.end < 0:
            raise ValueError("Step indices must be non-negative.")
        if self.end < self.start:
            raise ValueError("Step range end must be >= start.")

    @property
    def count(self) -> int:
        return self.end - self.start

    def __iter__(self) -> Iterator[int]:
        return iter(range(self.start, self.end))


def _check(total_steps: int, world_size: int, rank: int) -> None:
    if total_steps <= 0:
        raise ValueError("total_steps must be positive.")
    if world_size <= 0:
        raise ValueError("world_size must be positive.")
    if rank < 0 or rank >= world_size:
        raise ValueError("rank must satisfy 0 <= rank < world_size.")


def assign_steps(total_steps: int, world_size: int, rank: int) -> StepRange:
    """Equal contiguous split; raises ``ValueError`` when ``total_steps % world_size != 0``.

    Same contract as reference ``step_assignment.py:35-69``.
    """
    _check(total_steps, world_size, rank)
    per_stage, rem = divmod(total_steps, world_size)
    if rem:
        raise ValueError("total_steps must be divisible by world_size for uniform step assignment.")
    return StepRange(start=rank * per_stage, end=(rank + 1) * per_stage)


def stage_sizes(total_steps: int, world_size: int) -> List[int]:
    """Stage sizes of the uneven split: the first ``total_steps % world_size`` stages take one more."""
    _check(total_steps, world_size, 0)
    base, rem = divmod(total_steps, world_size)
    return [base + (1 if r < rem else 0) for r in range(world_size)]


def assign_steps_uneven(total_steps: int, world_size: int, rank: int) -> StepRange:
    """Contiguous split that tolerates remainders (extension; identical to ``assign_steps`` when
    divisible).  Stages may be empty when ``world_size > total_steps``."""
    _check(total_steps, world_size, rank)
    base, rem = divmod(total_steps, world_size)
    start = rank * base + min(rank, rem)
    return StepRange(start=start, end=start + base + (1 if rank < rem else 0))
