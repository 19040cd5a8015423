"""Batch sharding across the GPUs of one box: contiguous, independent image ranges, no data-path collective
(eval-mode BatchNorm uses running statistics and CBAM pools per sample, so per-image results are identical to the
unsharded forward — SURVEY 8e).  One process per GPU (torchrun); this module only decides who owns which images."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of rank's contiguous shard; the first n_items % world ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(max(0, n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_shards(n_items: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n_items, r, world) for r in range(world)]
