"""Batch sharding for multi-GPU runs: gates are independent, keys are replicated on
every GPU, a batch is split into contiguous slices — one per rank — and there is NO collective on the
data path."""
from __future__ import annotations


def shard_range(count: int, rank: int, world: int):
    """Contiguous slice [start, stop) of `count` gates owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(count, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
