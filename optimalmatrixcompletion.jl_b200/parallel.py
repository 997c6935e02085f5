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


class LibraryComm:
    """The engine's own exchange (csrc/omc_comm.cu: NCCL bound at run time inside libomc_b200.so), the calls a Julia host
    makes through the C ABI: omc_comm_unique_id / omc_comm_init / omc_allreduce_min / omc_allgather.  `bootstrap` ships the
    128-byte NCCL id from rank 0 to the other ranks by any host channel (here: a callable, e.g. a torch.distributed
    broadcast or a file); world = 1 needs nothing."""

    def __init__(self, rank: int, world: int, bootstrap=None):
        import ctypes as C
        from . import _lib
        self._C, self.lib, self.rank, self.world = C, _lib.load(), int(rank), int(world)
        idbuf = (C.c_uint8 * 128)()
        if world > 1:
            if rank == 0:
                self._check(self.lib.omc_comm_unique_id(idbuf))
            raw = bootstrap(bytes(idbuf) if rank == 0 else None)
            idbuf = (C.c_uint8 * 128).from_buffer_copy(raw)
        self._check(self.lib.omc_comm_init(self.rank, self.world, idbuf))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"libomc_b200 comm error {rc}: {self.lib.omc_comm_last_error().decode()}")

    def allreduce_min(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        self._check(self.lib.omc_allreduce_min(v.ctypes.data_as(self._C.POINTER(self._C.c_double)), v.size))
        return v

    def allgather(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros((self.world, v.size))
        self._check(self.lib.omc_allgather(v.ctypes.data_as(self._C.POINTER(self._C.c_double)), v.size,
                                           out.ctypes.data_as(self._C.POINTER(self._C.c_double))))
        return out

    def close(self):
        self.lib.omc_comm_destroy()


def lockstep_cost(costs_desc: Sequence[float], floor_nodes: int) -> float:
    """Predicted time of one lockstep relaxation of a shard (batched engine): iteration `it` costs max(active(it), floor_nodes)
    node-units -- below `floor_nodes` active nodes an iteration is bound by its launch latencies, not by the nodes' bytes.
    With iteration counts c_1 >= c_2 >= ... that is  floor_nodes * c_1 + sum_{j > floor_nodes} c_j."""
    if not costs_desc:
        return 0.0
    f = max(int(floor_nodes), 0)
    if f == 0:
        return float(sum(costs_desc))
    return f * float(costs_desc[0]) + float(sum(costs_desc[f:]))


def balanced_partition(costs: Sequence[float], world: int, floor_nodes: int = 0) -> List[List[int]]:
    """Frontier re-balancing plan: longest-predicted-first greedy (LPT) partition of node indices over `world` ranks by their
    predicted cost (ADMM iterations of the node's last relaxation, or of its parent).  Deterministic: every rank computes the
    same plan from the all-gathered costs, and takes its own part -- node descriptors (pool ids + direction codes) are
    replicated, so no node data moves.  Each part is ordered longest first.  floor_nodes > 0 uses the lockstep cost model
    (lockstep_cost): a shard holding a straggler gets fewer node-iterations, because the straggler's tail runs alone."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    parts: List[List[int]] = [[] for _ in range(world)]
    vals: List[List[float]] = [[] for _ in range(world)]
    loads = [0.0] * world
    for i in order:
        c = float(costs[i])
        r = min(range(world), key=lambda q: (lockstep_cost(vals[q] + [c], floor_nodes) if floor_nodes > 0 else loads[q], len(parts[q]), q))
        parts[r].append(i)
        vals[r].append(c)
        loads[r] += c
    return parts


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
