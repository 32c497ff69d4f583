"""Sharding of independent leaf-evaluation work across GPUs (one evaluator per GPU, no data-path collective).

The reference scales the same way: one self-play / eval process per GPU pinned with CUDA_VISIBLE_DEVICES
(``python/rl_loop/selfplay.py:51-64``, ``python/rl_loop/train_sp_eval.py:97-142``); the only cross-device step is
summing counters on the host.  These helpers are what ``bench.py`` uses under torchrun.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_items`` independent units (positions / games) owned by ``rank``.
    Sizes differ by at most one; every item is owned by exactly one rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def whole_job_rate(units_per_rank_per_step: int, steps: int, world: int, max_rank_seconds: float) -> float:
    """Units all ranks processed divided by the slowest rank's time (weak scaling: per-GPU work is fixed)."""
    return units_per_rank_per_step * steps * world / max_rank_seconds
