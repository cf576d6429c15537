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
