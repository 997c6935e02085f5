"""Python mirror of the Julia host loop (node queue, branching, cut bookkeeping, incumbent).

In production this logic stays in Julia (optimalmatrixcompletion.jl_b200/julia/OMCB200.jl); neither
Julia nor Mosek exists in this image, so the tests and the benchmark drive the C ABI from this
line-for-line mirror of /root/reference/src/OptimalMatrixCompletion.jl (OMC.jl).  Only bookkeeping
happens here; every numerical body is a libomc_b200.so call.
"""
import heapq
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .engine import Cut, Problem, LABELS, default_opts, smallest_eigvecs_batch


@dataclass
class BBNode:
    """OMC.jl:42-58 (U_lower/U_upper are constant on the disjunctive path and not carried)."""
    node_id: int
    parent_id: int
    LB: float
    depth: int
    master_feasible: bool = False
    disjunctive_cuts: List[Cut] = field(default_factory=list)
    warm_id: int = -1          # engine extension: state-pool record of the parent's ADMM state


def child_directions(cut_type: str, k: int):
    """OMC.jl:2479-2493: enumerate(Iterators.product(repeat([labels], k)...)), first factor fastest."""
    lab = LABELS[cut_type]
    P = len(lab)
    out = []
    for ind in range(P ** k):
        dirs, q = [], ind
        for _ in range(k):
            dirs.append(lab[q % P])
            q //= P
        out.append((ind + 1, dirs))
    return out


def create_matrix_cut_child_nodes(problem: Problem, node: BBNode, breakpoint_vec: np.ndarray, U: np.ndarray,
                                  counter: int, objective_relax: float, warm_id: int = -1) -> List[BBNode]:
    """Host part of OMC.jl:2411-2543: the breakpoint vector comes from the GPU separation oracle; children
    share the parent's cut list plus the new cut (x, U, directions); node_id = counter + ind (OMC.jl:2524)."""
    cid = problem.add_cut(breakpoint_vec, U)
    kids = []
    for ind, dirs in child_directions(problem.cut_type, problem.k):
        kids.append(BBNode(node_id=counter + ind, parent_id=node.node_id, LB=objective_relax, depth=node.depth + 1,
                           disjunctive_cuts=node.disjunctive_cuts + [Cut(cid, breakpoint_vec, U, dirs)],
                           warm_id=warm_id))
    return kids


def expand_frontier(problem: Problem, target: int, opts=None, nev: int = 1, cutoff: float = float("inf"),
                    max_levels: int = 64):
    """Breadth-first expansion from the root, one frontier batch per level, until `target` open nodes exist
    (the frontier workload).  Children of nodes whose bound exceeds `cutoff`, of
    infeasible nodes and of master-feasible nodes are not created (OMC.jl:777-800, 814-837).
    Returns (open_nodes sorted by node_id, stats)."""
    opts = opts or default_opts()
    opts.cutoff = cutoff
    root = BBNode(node_id=1, parent_id=0, LB=-np.inf, depth=0)
    level, counter, relaxed, total_iters = [root], 1, 0, 0
    for _ in range(max_levels):
        if len(level) >= target or not level:
            break
        res = problem.relax_batch([nd.disjunctive_cuts for nd in level], opts)
        relaxed += len(level)
        total_iters += sum(r["iters"] for r in res)
        keep = [i for i, r in enumerate(res) if r["feasible"] and r["status_code"] in (0, 1) and r["objective"] <= cutoff]
        nxt: List[BBNode] = []
        if keep:
            Y = np.stack([res[i]["Y"] for i in keep]); U = np.stack([res[i]["U"] for i in keep])
            lam, vec, bp, feas = smallest_eigvecs_batch(Y, U, nev)
            for q, i in enumerate(keep):
                level[i].LB = res[i]["objective"]
                if feas[q]:
                    level[i].master_feasible = True
                    continue
                kids = create_matrix_cut_child_nodes(problem, level[i], bp[q], res[i]["U"], counter, res[i]["objective"])
                counter += len(kids)
                nxt.extend(kids)
        level = nxt
    level.sort(key=lambda nd: nd.node_id)
    return level[:target], dict(relaxed=relaxed, total_iters=total_iters, counter=counter)


def shard_block_cyclic(items: list, rank: int, world: int) -> list:
    """Frontier sharding across GPUs: open nodes are independent, node i -> rank i mod world."""
    return items[rank::world]
