"""Sample-range data parallelism: every sample is independent, so a take() or a
batch is split into contiguous global ranges, one per rank; each GPU holds a replica of the curve
and writes its own output slice. No collective is needed on the data path."""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous global range [lo, hi) owned by `rank` of `world`: lo = rank*n/world (floor)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def shard_sizes(n_total: int, world: int) -> list[int]:
    return [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
