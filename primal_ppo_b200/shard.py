"""Sharding of worlds over ranks (one process per GPU).  Worlds are independent, so the env path has no collective:
rank r owns the contiguous range ``shard_range(W, r, G)`` and keys its Philox draws with the global
world index (``world_offset``), so a world behaves identically on whichever rank owns it."""
from __future__ import annotations

from .scenario import Scenario


def shard_range(num_worlds: int, rank: int, world_size: int):
    """Contiguous, balanced: the first ``num_worlds % world_size`` ranks get one extra world."""
    base, extra = divmod(num_worlds, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_scenario(sc: Scenario, rank: int, world_size: int):
    """Returns (scenario slice of this rank, world_offset to pass to BatchedMapfGym)."""
    lo, hi = shard_range(sc.num_worlds, rank, world_size)
    return sc.slice(lo, hi), lo
