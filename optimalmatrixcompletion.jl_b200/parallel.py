"""Frontier sharding across GPUs.

Open branch-and-bound nodes are independent given (A, indices, gamma) and their own cut lists
(OMC.jl:747-754 reads only `current_node` + globals), so the popped batch is partitioned block-cyclically,
one process per GPU, with the problem and the cut pool replicated and NO data-path collective.  The only
exchange is the tiny all-reduce-min of [incumbent upper bound, smallest open lower bound] after a batch,
which goes over NCCL (NVLink) on the GPU box and over gloo in the CPU tests.
"""
from typing import List, Sequence, Tuple

import numpy as np


def shard_block_cyclic(items: Sequence, rank: int, world: int) -> list:
    return list(items[rank::world])


def unshard_block_cyclic(parts: List[list]) -> list:
    """Inverse of shard_block_cyclic: interleaves the per-rank result lists back into pop order."""
    world = len(parts)
    total = sum(len(p) for p in parts)
    out = [None] * total
    for r, p in enumerate(parts):
        out[r::world] = p
    return out


def allreduce_bounds(upper: float, lower: float, device=None) -> Tuple[float, float]:
    """all-reduce-min of [incumbent, min open lower bound] over the default process group (2 x f64)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([upper, lower], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t[0]), float(t[1])


def rebalance_counts(iters_per_rank: Sequence[int], nodes_per_rank: Sequence[int]) -> np.ndarray:
    """Frontier rebalancing plan: given each rank's measured ADMM iterations for its last shard, returns how many
    nodes of the next batch each rank should take so that predicted work (iterations per node x nodes) evens out."""
    it = np.maximum(np.asarray(iters_per_rank, dtype=float), 1.0)
    nd = np.maximum(np.asarray(nodes_per_rank, dtype=float), 1.0)
    speed = nd / it                                   # nodes per unit of work
    share = speed / speed.sum()
    total = int(np.sum(nodes_per_rank))
    counts = np.floor(share * total).astype(int)
    counts[: total - counts.sum()] += 1
    return counts
