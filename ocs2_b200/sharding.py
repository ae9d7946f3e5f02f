"""Sharding of a batch of independent LQ problems over the GPUs of one box.

Problems share no data and the path has no exchange step, so the only multi-GPU logic is index arithmetic: contiguous blocks
``[begin, begin + count)`` of the global problem index per rank (one process per GPU), host scatter of the inputs and gather of the
outputs. No collective runs on the data path; ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests) is used only to
gather results / timings when the caller wants them in one place.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np


def shard_bounds(num_problems: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: the first `num_problems % world_size` ranks own one problem more."""
    if world_size < 1 or not (0 <= rank < world_size) or num_problems < 0:
        raise ValueError("invalid shard request")
    base, extra = divmod(num_problems, world_size)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def all_shard_bounds(num_problems: int, world_size: int) -> List[Tuple[int, int]]:
    return [shard_bounds(num_problems, world_size, r) for r in range(world_size)]


def shard_arrays(arrays: Dict[str, Optional[np.ndarray]], world_size: int, rank: int, shared=("time",)) -> Dict[str, Optional[np.ndarray]]:
    """Slices every per-problem array (leading dimension = problem) to this rank's block; `shared` keys are passed through."""
    lead = {v.shape[0] for k, v in arrays.items() if v is not None and k not in shared}
    if len(lead) != 1:
        raise ValueError(f"per-problem arrays disagree on the batch size: {sorted(lead)}")
    begin, count = shard_bounds(lead.pop(), world_size, rank)
    return {k: (v if (v is None or k in shared) else v[begin:begin + count]) for k, v in arrays.items()}


def gather_arrays(local: Dict[str, np.ndarray], num_problems: int, dist=None, dst: int = 0) -> Optional[Dict[str, np.ndarray]]:
    """Reassembles per-problem result arrays of all ranks in global problem order on rank `dst` (None elsewhere).
    `dist` is an initialised torch.distributed module (any backend with gather_object); None means single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    begin, count = shard_bounds(num_problems, world, rank)
    for k, v in local.items():
        if v.shape[0] != count:
            raise ValueError(f"{k}: rank {rank} holds {v.shape[0]} problems, its shard has {count}")
    gathered = [None] * world if rank == dst else None
    dist.gather_object((begin, local), gathered, dst=dst)
    if rank != dst:
        return None
    gathered.sort(key=lambda t: t[0])
    return {k: np.concatenate([g[1][k] for g in gathered], axis=0) for k in local}
