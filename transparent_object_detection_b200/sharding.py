"""Batch sharding across the GPUs of one box: independent images, one process per GPU, a full
weight replica each, NO collective on the data path.  The only exchange is a host-side gather of the per-rank
detection lists (reference output type: list of None | float32 (n, 6)), concatenated in rank order so that the result
equals the single-GPU result exactly."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image range [lo, hi) of `rank`: [r*N/W, (r+1)*N/W) (integer arithmetic; ranges tile [0, N))."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def weighted_shard_ranges(n_items: int, weights: Sequence[float], granule: int = 1) -> List[Tuple[int, int]]:
    """Contiguous ranges [lo, hi) per rank, in rank order, whose sizes follow `weights` (each rank's capacity in items/s:
    min(its compute rate, its host->device bandwidth / bytes per item)) in multiples of `granule` items (a batch), largest
    remainders first; the ranges tile [0, n_items).  On a box whose GPUs see different host bandwidth (measured on the
    8 x B200 guest: 20.4 vs 35.8 GB/s, profiles/r2_config3_and_h2d_ceiling.md) equal ranges make every rank wait for the
    slowest upload; the result is the same list of detections either way (every op is per-image)."""
    world = len(weights)
    if world == 0 or granule <= 0 or n_items % granule != 0:
        raise ValueError(f"{n_items} items cannot be split into granules of {granule} over {world} ranks")
    if any((not np.isfinite(w)) or w < 0 for w in weights) or sum(weights) <= 0:
        raise ValueError(f"bad capacities {list(weights)}")
    units = n_items // granule
    ideal = [units * float(w) / float(sum(weights)) for w in weights]
    take = [int(np.floor(x)) for x in ideal]
    for r in sorted(range(world), key=lambda r_: (-(ideal[r_] - take[r_]), r_))[:units - sum(take)]:
        take[r] += 1
    out, lo = [], 0
    for t in take:
        out.append((lo, lo + t * granule))
        lo += t * granule
    assert lo == n_items
    return out


def gather_detections(local: Sequence[Optional[np.ndarray]], group=None, dst: int = 0) -> Optional[List[Optional[np.ndarray]]]:
    """Host-side gather of per-rank detection lists to `dst` (rank order).  Returns the concatenated list on `dst`,
    None elsewhere.  Works on any torch.distributed backend (object gather goes through host memory)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return list(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(list(local), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[Optional[np.ndarray]] = []
    for part in bucket:
        out.extend(part)
    return out
