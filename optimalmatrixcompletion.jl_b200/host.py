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
    Shor_info: Optional[object] = None   # BBNodeShorInfo (OMC.jl:37-40): shared object of (constraints_indexes, SOC_constraints_indexes), 0-based


@dataclass
class BBNodeShorInfo:
    """OMC.jl:37-40, 0-based: constraints_indexes (N, 4) int32 minors, SOC_constraints_indexes (M, 2) int32 coordinates."""
    constraints_indexes: np.ndarray
    SOC_constraints_indexes: np.ndarray


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


from .parallel import shard_block_cyclic  # noqa: E402,F401


# ==================================================================================================
# Tree bookkeeping (OMC.jl:1149-1244) and the main loop (OMC.jl:140-1146)
# ==================================================================================================

class JuliaPriorityQueue:
    """Binary min-heap with the update rules of DataStructures.jl 0.18 PriorityQueue (the un-vendored
    dependency behind ``tree.lower_bounds``, Manifest.toml), restated so that ties -- children are
    enqueued with their parent's objective (OMC.jl:1201) -- pop in the same order."""

    def __init__(self, pairs=()):
        self.xs = [list(p) for p in pairs]
        self.index = {p[0]: i for i, p in enumerate(self.xs)}
        for i in range(len(self.xs) // 2 - 1, -1, -1):
            self._down(i)

    def __len__(self):
        return len(self.xs)

    def _down(self, i):
        x = self.xs[i]
        n = len(self.xs)
        while 2 * i + 1 < n:
            l, r = 2 * i + 1, 2 * i + 2
            j = l if (r >= n or self.xs[l][1] < self.xs[r][1]) else r
            if self.xs[j][1] < x[1]:
                self.index[self.xs[j][0]] = i
                self.xs[i] = self.xs[j]
                i = j
            else:
                break
        self.index[x[0]] = i
        self.xs[i] = x

    def _up(self, i, force=False):
        x = self.xs[i]
        while i > 0:
            j = (i - 1) // 2
            if force or x[1] < self.xs[j][1]:
                self.index[self.xs[j][0]] = i
                self.xs[i] = self.xs[j]
                i = j
            else:
                break
        self.index[x[0]] = i
        self.xs[i] = x

    def enqueue(self, key, value):
        self.xs.append([key, value])
        self._up(len(self.xs) - 1)

    def dequeue_pair(self):
        x = self.xs[0]
        y = self.xs.pop()
        if self.xs:
            self.xs[0] = y
            self.index[y[0]] = 0
            self._down(0)
        del self.index[x[0]]
        return x[0], x[1]

    def delete(self, key):
        self._up(self.index[key], force=True)
        self.dequeue_pair()

    def peek(self):
        return self.xs[0][0], self.xs[0][1]

    def items(self):
        return [(a, b) for a, b in self.xs]


@dataclass
class BBTree:
    """OMC.jl:60-71."""
    nodes: Dict[int, BBNode]
    node_ids: List[int]
    counter: int
    last_updated_counter: int
    nodes_explored: int
    nodes_remaining: int
    best_upper_bound: float
    best_lower_bound: float
    now_gap: float
    lower_bounds: JuliaPriorityQueue


def retrieve_node_from_tree(tree: BBTree, node_selection_here: str) -> BBNode:
    """OMC.jl:1164-1182."""
    if node_selection_here == "breadthfirst":
        nid = tree.node_ids.pop(0)
        tree.lower_bounds.delete(nid)
    elif node_selection_here == "bestfirst":
        nid, _ = tree.lower_bounds.dequeue_pair()
        tree.node_ids.remove(nid)
    elif node_selection_here == "depthfirst":
        nid = tree.node_ids.pop()
        tree.lower_bounds.delete(nid)
    else:
        raise ValueError(node_selection_here)
    node = tree.nodes.pop(nid)
    tree.nodes_explored += 1
    tree.nodes_remaining -= 1
    return node


def add_nodes_to_tree(tree: BBTree, child_nodes: List[BBNode], parent_objective: float):
    """OMC.jl:1185-1205."""
    for i, nd in enumerate(child_nodes, start=1):
        tree.nodes[tree.counter + i] = nd
    new_ids = list(range(tree.counter + 1, tree.counter + len(child_nodes) + 1))
    tree.node_ids.extend(new_ids)
    for nid in new_ids:
        tree.lower_bounds.enqueue(nid, parent_objective)
    tree.counter += len(child_nodes)
    tree.nodes_remaining += len(child_nodes)


def update_tree_lower_bounds(tree: BBTree, pending_min: float = float("inf")) -> bool:
    """OMC.jl:1207-1218.  pending_min: smallest bound among nodes of the current batch that were popped but
    not consumed yet (they left the queue early); +inf when frontier_batch = 1, i.e. the reference's rule."""
    if len(tree.lower_bounds) == 0 and pending_min == float("inf"):
        return True
    minval = pending_min
    if len(tree.lower_bounds):
        minval = min(minval, tree.lower_bounds.peek()[1])
    if minval > tree.best_lower_bound:
        tree.best_lower_bound = minval
        return True
    return False


def prune_dominated_nodes(tree: BBTree):
    """OMC.jl:1220-1244 (rebuilds the queue in dequeue order, exactly like the reference)."""
    new_lb = JuliaPriorityQueue()
    removed = []
    while len(tree.lower_bounds):
        nid, lb = tree.lower_bounds.dequeue_pair()
        if lb > tree.best_upper_bound:
            removed.append(nid)
            removed.extend(k for k, _ in tree.lower_bounds.items())
            break
        new_lb.enqueue(nid, lb)
    tree.lower_bounds = new_lb
    dropped = []
    for nid in removed:
        nd = tree.nodes.pop(nid, None)
        if nd is not None:
            dropped.append(nd)
    if removed:
        rs = set(removed)
        tree.node_ids = [i for i in tree.node_ids if i not in rs]
    tree.nodes_remaining = len(tree.nodes)
    return dropped


def compute_gap(lower: float, upper: float) -> float:
    """OMC.jl:173-179."""
    return float("inf") if lower < 0 else upper / lower - 1.0


def matrix_completion_branchandbound(k: int, A: np.ndarray, indices: np.ndarray, gamma: float, *,
                                     node_selection: str = "breadthfirst", bestfirst_depthfirst_cutoff: int = 10000,
                                     gap: float = 1e-4, use_disjunctive_cuts: bool = True,
                                     disjunctive_cuts_type: Optional[str] = None,
                                     disjunctive_cuts_breakpoints: Optional[str] = None,
                                     add_Shor_valid_inequalities: bool = False,
                                     Shor_valid_inequalities_noisy_rank1_num_entries_present=(1, 2, 3, 4),
                                     add_Shor_valid_inequalities_fraction: float = 1.0,
                                     add_Shor_valid_inequalities_iterative: bool = False,
                                     max_update_Shor_indices_probability: float = 1.0,
                                     min_update_Shor_indices_probability: float = 0.1,
                                     update_Shor_indices_probability_decay_rate: float = 1.1,
                                     update_Shor_indices_n_minors: int = 100, root_only: bool = False,
                                     altmin_flag: bool = True, max_altmin_probability: float = 1.0,
                                     min_altmin_probability: float = 0.005, altmin_probability_decay_rate: float = 1.1,
                                     altmin_root_n_iters: int = 1, use_max_steps: bool = False, max_steps: int = 1000000,
                                     time_limit: int = 3600, update_step: int = 1000, verbosity: int = 0,
                                     # engine knobs (defaults leave the reference's behaviour unchanged)
                                     frontier_batch: int = 1, relax_opts=None, use_cutoff: bool = True,
                                     warm_start: bool = False, device: Optional[int] = None, stop_at_open_nodes: int = 0, seed: int = 0):
    """Mirror of OMC.jl:140-1146 (disjunctive path).  Returns (solution, printlist, instance) with the
    reference's keys.  frontier_batch = 1 reproduces the reference's sequence of pops; larger batches pop B
    nodes, relax them in one launch and consume the results in pop order.
    stop_at_open_nodes > 0 stops as soon as that many nodes are open and returns them in instance["open_nodes"]
    (used to build benchmark frontiers)."""
    from .engine import alternating_minimization as _altmin
    if not use_disjunctive_cuts:
        raise NotImplementedError("use_disjunctive_cuts = false (McCormick path) is out of scope (SURVEY.md section 2)")
    if add_Shor_valid_inequalities:
        if warm_start:
            raise ValueError("warm_start needs the persistent engine; Shor rows run on the batched engine, which starts every node cold")
        if not 0.0 <= add_Shor_valid_inequalities_fraction <= 1.0:                                            # OMC.jl:256-263
            raise ValueError(f"Argument `add_Shor_valid_inequalities_fraction` = {add_Shor_valid_inequalities_fraction} out of bounds [0.0, 1.0].")
    else:
        add_Shor_valid_inequalities_fraction = None                                                           # OMC.jl:264-266
    if add_Shor_valid_inequalities and add_Shor_valid_inequalities_iterative:                                 # OMC.jl:296-324
        if not 0.0 <= max_update_Shor_indices_probability <= 1.0:
            raise ValueError(f"Argument `max_update_Shor_indices_probability` = {max_update_Shor_indices_probability} out of bounds [0.0, 1.0].")
        if not 0.0 < min_update_Shor_indices_probability < 1.0:
            raise ValueError(f"Argument `min_update_Shor_indices_probability` = {min_update_Shor_indices_probability} out of bounds (0.0, 1.0).")
        if not 1.0 < update_Shor_indices_probability_decay_rate:
            raise ValueError(f"Argument `update_Shor_indices_probability_decay_rate` = {update_Shor_indices_probability_decay_rate} out of bounds (1.0, ∞).")
        if not 1.0 <= update_Shor_indices_n_minors:
            raise ValueError(f"Argument `update_Shor_indices_n_minors` = {update_Shor_indices_n_minors} out of bounds [1.0, ∞).")
    else:
        max_update_Shor_indices_probability = min_update_Shor_indices_probability = None                     # OMC.jl:325-330
        update_Shor_indices_probability_decay_rate = update_Shor_indices_n_minors = None
    if disjunctive_cuts_type not in ("linear", "linear2", "linear3"):
        raise ValueError('Invalid input for disjunctive cuts type.\nDisjunctive cuts type must be either "linear" or '
                         f'"linear2" or "linear3";\n{disjunctive_cuts_type} supplied instead.')         # OMC.jl:218-224
    if disjunctive_cuts_breakpoints not in ("smallest_1_eigvec", "smallest_2_eigvec"):
        raise ValueError('Invalid input for disjunctive cuts breakpoints.\nDisjunctive cuts type must be either '
                         f'"smallest_1_eigvec" or "smallest_2_eigvec";\n{disjunctive_cuts_breakpoints} supplied instead.')
    if node_selection not in ("breadthfirst", "bestfirst", "depthfirst", "bestfirst_depthfirst"):
        raise ValueError(f"Invalid input for node selection. {node_selection} supplied instead.")            # OMC.jl:233-238
    if A.shape != indices.shape:
        raise ValueError("Dimension mismatch. \nInput matrix A must have size (n, m);\nInput matrix indices must have size (n, m).")
    n, m = A.shape
    if not n <= m:
        raise ValueError(f"Input matrix A must have size (n, m) with n <= m.\nCurrent size is {A.shape}.")  # OMC.jl:249-254
    if altmin_flag:
        if not 0.0 <= max_altmin_probability <= 1.0:
            raise ValueError("Argument `max_altmin_probability` out of bounds [0.0, 1.0].")
        if not 0.0 < min_altmin_probability < 1.0:
            raise ValueError("Argument `min_altmin_probability` out of bounds (0.0, 1.0).")
        if not 1.0 < altmin_probability_decay_rate:
            raise ValueError("Argument `altmin_probability_decay_rate` out of bounds (1.0, inf).")
    nev = 1 if disjunctive_cuts_breakpoints == "smallest_1_eigvec" else 2
    rng = np.random.default_rng(seed)            # Random.seed!(0), OMC.jl:333 (Julia's stream is not reproducible here)
    printlist: List[str] = []

    def add_message(msgs):
        for s in msgs:
            if verbosity >= 1:
                print(s, end="")
            printlist.append(s)

    start_time = time.time()
    pool = max(4, 4 * frontier_batch) if warm_start else 0
    problem = Problem(k, A, indices, gamma, disjunctive_cuts_type, state_pool_capacity=pool, device=device)
    if add_Shor_valid_inequalities:
        # OMC.jl:646-669: all 2 x 2 minors with the listed numbers of observed entries, a random fraction of them, and an RSOC row
        # on every coordinate no kept minor covers
        from .engine import shor_constraint_indexes as _shor_idx, generate_violated_Shor_minors as _violated
        minors, soc = _shor_idx(problem, list(Shor_valid_inequalities_noisy_rank1_num_entries_present), with_soc=True)
        shor_candidates = minors                     # all minors of the pattern: depends on the mask only (OMC.jl:2621-2624)
        if add_Shor_valid_inequalities_iterative:
            # OMC.jl:670-675: no minors at the root, an RSOC row on every coordinate (column-major order of Iterators.product)
            minors = np.zeros((0, 4), np.int32)
            soc = np.array([(i, j) for j in range(m) for i in range(n)], np.int32)
        elif add_Shor_valid_inequalities_fraction < 1.0:
            minors = minors[rng.random(len(minors)) < add_Shor_valid_inequalities_fraction]                   # randsubseq
            cov = np.zeros((n, m), bool)
            cov[minors[:, 0], minors[:, 2]] = cov[minors[:, 0], minors[:, 3]] = True
            cov[minors[:, 1], minors[:, 2]] = cov[minors[:, 1], minors[:, 3]] = True
            soc = np.argwhere(~cov).astype(np.int32)
        root_shor = BBNodeShorInfo(minors, soc)
        problem.set_shor(minors, soc)
        shor_loaded = [root_shor]                    # the structure currently attached to the problem
        instance_shor = dict(constraints_indexes=minors, SOC_constraints_indexes=soc)
    solve_time_altmin = solve_time_relaxation = 0.0
    dict_solve_times_altmin, dict_num_iterations_altmin, dict_solve_times_relaxation = [], [], []
    cnt = dict(nodes_dominated=0, nodes_relax_infeasible=0, nodes_relax_feasible=0, nodes_relax_feasible_pruned=0,
               nodes_master_feasible=0, nodes_master_feasible_improvement=0, nodes_relax_feasible_split=0,
               nodes_relax_feasible_split_altmin=0, nodes_relax_feasible_split_altmin_improvement=0)
    instance = {"run_log": [], "run_details": {}}

    # ---- root heuristic (OMC.jl:521-621): host-side LAPACK svd, GPU alt-min / objective / MSE
    A0 = np.where(indices, A, 0.0)
    U_base = np.linalg.svd(A0)[0][:, :k]
    sc = np.abs(U_base).max()
    objective_initial, X_initial = np.inf, None
    for it in range(1, altmin_root_n_iters + 1):
        U_init = U_base if it == 1 else U_base + sc * rng.standard_normal((n, k))
        am = _altmin(problem, U_init, [], time_limit=time_limit)
        solve_time_altmin += am["solve_time"]
        dict_solve_times_altmin.append((0, 0, am["solve_time"]))
        X_ = am["U"] @ am["V"]
        obj_ = problem.objective_mse(X_)[0]
        if obj_ < objective_initial:
            objective_initial, X_initial = obj_, X_
    U_initial = np.linalg.svd(X_initial)[0][:, :k]
    o4 = problem.objective_mse(X_initial)
    solution = {"objective_initial": o4[0], "MSE_in_initial": o4[1], "MSE_out_initial": o4[2], "MSE_all_initial": o4[3],
                "Y_initial": U_initial @ U_initial.T, "U_initial": U_initial, "X_initial": X_initial,
                "objective_initial_time_found": time.time() - start_time,
                "objective": o4[0], "objective_time_found": time.time() - start_time,
                "MSE_in": o4[1], "MSE_out": o4[2], "MSE_all": o4[3],
                "Y": U_initial @ U_initial.T, "U": U_initial, "X": X_initial}
    objective_initial = o4[0]

    root = BBNode(node_id=1, parent_id=0, LB=-np.inf, depth=0, Shor_info=root_shor if add_Shor_valid_inequalities else None)
    tree = BBTree(nodes={1: root}, node_ids=[1], counter=1, last_updated_counter=1, nodes_explored=0, nodes_remaining=1,
                  best_upper_bound=objective_initial, best_lower_bound=-np.inf, now_gap=np.inf,
                  lower_bounds=JuliaPriorityQueue([(1, np.inf)]))

    def add_update(altmin=False, print_message=True):
        tree.now_gap = compute_gap(tree.best_lower_bound, tree.best_upper_bound)
        msg = "| %10d | %10d | %10d | %10f | %10f | %10f | %10.3f  s  |" % (
            tree.nodes_explored, tree.counter, tree.nodes_remaining, tree.best_lower_bound, tree.best_upper_bound,
            tree.now_gap, time.time() - start_time) + (" - A\n" if altmin else "\n")
        if print_message:
            add_message([msg])
        instance["run_log"].append((tree.nodes_explored, tree.counter, tree.nodes_remaining, tree.best_lower_bound,
                                    tree.best_upper_bound, tree.now_gap, time.time() - start_time))
        tree.last_updated_counter = tree.counter

    opts = relax_opts or default_opts()
    free_states = list(range(pool))
    state_refs: Dict[int, int] = {}
    pool_exhausted = 0

    def release_parent_state(nd: BBNode):
        """A child that leaves the tree (relaxed, dominated or pruned) gives back its share of the parent's warm-start record."""
        if nd.warm_id >= 0 and nd.warm_id in state_refs:
            state_refs[nd.warm_id] -= 1
            if state_refs[nd.warm_id] == 0:
                free_states.append(nd.warm_id)
                del state_refs[nd.warm_id]
        nd.warm_id = -1
    root_node_timeout = False
    while (tree.now_gap > gap and not (use_max_steps and tree.counter >= max_steps)
           and time.time() - start_time <= time_limit):
        if len(tree.nodes) == 0:
            break
        if stop_at_open_nodes and len(tree.nodes) >= stop_at_open_nodes:
            break
        # ---- pop up to frontier_batch nodes (OMC.jl:709-728)
        batch: List[BBNode] = []
        while len(batch) < frontier_batch and len(tree.nodes) > 0:
            sel = node_selection
            if node_selection == "bestfirst_depthfirst":
                sel = "depthfirst" if len(tree.nodes) > bestfirst_depthfirst_cutoff else "bestfirst"
            batch.append(retrieve_node_from_tree(tree, sel))
        live = [nd for nd in batch if not nd.LB > tree.best_upper_bound]
        results = {}
        if live:
            opts.cutoff = tree.best_upper_bound if use_cutoff else float("inf")
            opts.time_limit_s = max(1.0, time_limit - (time.time() - start_time))
            save = None
            if warm_start:
                save = [free_states.pop() if free_states else -1 for _ in live]
                pool_exhausted += sum(1 for sid in save if sid < 0)
            if add_Shor_valid_inequalities and add_Shor_valid_inequalities_iterative:
                # the engine holds ONE Shor row structure per problem: nodes are relaxed in groups that share their BBNodeShorInfo
                # (all children of a split do, OMC.jl:2532-2539), the structure is swapped between groups (omc_problem_set_shor)
                res = [None] * len(live)
                groups = {}
                for q, nd in enumerate(live):
                    groups.setdefault(id(nd.Shor_info), []).append(q)
                for qs in groups.values():
                    info = live[qs[0]].Shor_info
                    if shor_loaded[0] is not info:
                        problem.set_shor(info.constraints_indexes, info.SOC_constraints_indexes)
                        shor_loaded[0] = info
                    for q, r in zip(qs, problem.relax_batch([live[q].disjunctive_cuts for q in qs], opts)):
                        res[q] = r
            else:
                res = problem.relax_batch([nd.disjunctive_cuts for nd in live], opts,
                                          warm_ids=[nd.warm_id for nd in live] if warm_start else None, save_ids=save)
            for q, (nd, r) in enumerate(zip(live, res)):
                r["save_id"] = save[q] if save else -1
                results[nd.node_id] = r
        if warm_start:                           # parents' records are no longer needed once their children left the queue
            for nd in batch:
                release_parent_state(nd)
        # separation oracle for the whole batch in one launch (OMC.jl:814, 2466-2477)
        eig = {}
        cand = [nd for nd in live if results[nd.node_id]["feasible"]]
        if cand:
            lam, vec, bp, feas = smallest_eigvecs_batch(np.stack([results[nd.node_id]["Y"] for nd in cand]),
                                                        np.stack([results[nd.node_id]["U"] for nd in cand]), nev)
            for q, nd in enumerate(cand):
                eig[nd.node_id] = (bp[q], bool(feas[q]))
        # ---- consume the results in pop order (OMC.jl:721-1073)
        for pos, current_node in enumerate(batch):
            pending_min = min([nd.LB for nd in batch[pos + 1:] if not nd.LB > tree.best_upper_bound] + [float("inf")])
            split_flag = True
            relax_result = None
            if current_node.LB > tree.best_upper_bound:
                split_flag = False
                cnt["nodes_dominated"] += 1
            if split_flag:
                relax_result = results[current_node.node_id]
                solve_time_relaxation += relax_result["solve_time"]
                dict_solve_times_relaxation.append((current_node.node_id, current_node.depth, relax_result["solve_time"]))
                if current_node.node_id == 1:
                    root_node_timeout = relax_result["termination_status"] == "TIME_LIMIT"
                if not relax_result["feasible"]:
                    cnt["nodes_relax_infeasible"] += 1
                    split_flag = False
                else:
                    cnt["nodes_relax_feasible"] += 1
                    objective_relax = relax_result["objective"]
                    if relax_result["termination_status"] != "OPTIMAL":
                        # SLOW_PROGRESS / TIME_LIMIT: the primal objective of an unconverged first-order iterate is no bound.
                        # The node keeps the larger of its inherited bound and the kernel's certified lower bound
                        # (dual objective - ||r_d||_inf ||w*||_1), so pruning and the children's
                        # bounds never rest on an uncertified value (the reference trusts Mosek's SLOW_PROGRESS point,
                        # OMC.jl:1871-1877, which is an interior-point iterate at ~1e-8).
                        cert = min(objective_relax, relax_result["lower_bound"])
                        objective_relax = max(current_node.LB, cert) if np.isfinite(current_node.LB) else cert
                    current_node.LB = objective_relax
                    if current_node.node_id == 1:
                        tree.best_lower_bound = objective_relax
                    if objective_relax > tree.best_upper_bound:
                        cnt["nodes_relax_feasible_pruned"] += 1
                        split_flag = False
            if split_flag and relax_result["termination_status"] == "OPTIMAL":
                if eig[current_node.node_id][1]:                      # matrix_completion_master_feasible, OMC.jl:814
                    current_node.master_feasible = True
                    cnt["nodes_master_feasible"] += 1
                    if objective_relax < tree.best_upper_bound:
                        cnt["nodes_master_feasible_improvement"] += 1
                        solution.update(objective=objective_relax, objective_time_found=time.time() - start_time,
                                        Y=relax_result["Y"].copy(), U=relax_result["U"].copy(), X=relax_result["X"].copy())
                        tree.best_upper_bound = objective_relax
                        add_update(print_message=verbosity >= 1)
                    split_flag = False
            elif split_flag and relax_result["termination_status"] == "TIME_LIMIT":
                add_update(print_message=verbosity >= 1)
                add_message(["Time limit reached.\n"])
            # alternating minimisation heuristic (OMC.jl:856-949): rand() is consumed once per processed node
            altmin_flag_now = False
            if altmin_flag:
                lim = np.log(max_altmin_probability / min_altmin_probability) / np.log(altmin_probability_decay_rate)
                prob = (min_altmin_probability if current_node.depth > lim
                        else max_altmin_probability / altmin_probability_decay_rate ** current_node.depth)
                altmin_flag_now = rng.random() < prob
            if split_flag and altmin_flag_now:
                U_rounded = np.linalg.svd(relax_result["Y"])[0][:, :k]                     # OMC.jl:873 (host LAPACK)
                am = _altmin(problem, U_rounded, current_node.disjunctive_cuts, time_limit=time_limit)
                cnt["nodes_relax_feasible_split_altmin"] += 1
                solve_time_altmin += am["solve_time"]
                dict_solve_times_altmin.append((current_node.node_id, current_node.depth, am["solve_time"]))
                dict_num_iterations_altmin.append((current_node.node_id, current_node.depth, am["n_iters"]))
                if am["converged"]:                                                        # OMC.jl:919
                    X_local = am["U"] @ am["V"]
                    U_local = np.linalg.svd(X_local)[0][:, :k]
                    objective_local = problem.objective_mse(X_local)[0]
                    if objective_local < tree.best_upper_bound:
                        cnt["nodes_relax_feasible_split_altmin_improvement"] += 1
                        solution.update(objective=objective_local, objective_time_found=time.time() - start_time,
                                        Y=U_local @ U_local.T, U=U_local, X=X_local)
                        tree.best_upper_bound = objective_local
                        add_update(altmin=True, print_message=verbosity >= 1)
            if split_flag:
                cnt["nodes_relax_feasible_split"] += 1
                sid = relax_result.get("save_id", -1)
                kids = create_matrix_cut_child_nodes(problem, current_node, eig[current_node.node_id][0], relax_result["U"],
                                                     tree.counter, objective_relax, warm_id=sid)
                if sid >= 0:
                    state_refs[sid] = len(kids)
                if add_Shor_valid_inequalities:
                    info = current_node.Shor_info
                    if add_Shor_valid_inequalities_iterative:
                        # OMC.jl:956-969: the deeper the node, the rarer the update; OMC.jl:2495-2518: the n_minors most violated
                        # minors of X join the node's list, their coordinates leave the RSOC list; all children share the result
                        lim = np.log(max_update_Shor_indices_probability / min_update_Shor_indices_probability) / np.log(update_Shor_indices_probability_decay_rate)
                        prob = (min_update_Shor_indices_probability if current_node.depth > lim
                                else max_update_Shor_indices_probability / update_Shor_indices_probability_decay_rate ** current_node.depth)
                        if rng.random() < prob:
                            _, new_minors = _violated(problem, relax_result["X"][None], shor_candidates, info.constraints_indexes,
                                                      int(update_Shor_indices_n_minors))
                            allm = np.concatenate([info.constraints_indexes, new_minors]).astype(np.int32)      # union: new ones are not in the old list
                            cov = np.zeros((n, m), bool)
                            cov[allm[:, 0], allm[:, 2]] = cov[allm[:, 0], allm[:, 3]] = True
                            cov[allm[:, 1], allm[:, 2]] = cov[allm[:, 1], allm[:, 3]] = True
                            old_soc = info.SOC_constraints_indexes
                            info = BBNodeShorInfo(allm, old_soc[~cov[old_soc[:, 0], old_soc[:, 1]]])                 # setdiff keeps the order
                            cnt["Shor_indices_updates"] = cnt.get("Shor_indices_updates", 0) + 1
                    for kd in kids:
                        kd.Shor_info = info
                add_nodes_to_tree(tree, kids, objective_relax)
            elif warm_start and relax_result is not None and relax_result.get("save_id", -1) >= 0:
                free_states.append(relax_result["save_id"])
            elif warm_start and relax_result is None and current_node.node_id in results:
                sid = results[current_node.node_id].get("save_id", -1)   # relaxed in this batch but dominated before its turn
                if sid >= 0:
                    free_states.append(sid)
            for nd in prune_dominated_nodes(tree):
                release_parent_state(nd)
            lower_bounds_updated = update_tree_lower_bounds(tree, pending_min)
            important = (lower_bounds_updated or current_node.node_id == 1
                         or tree.counter // update_step > tree.last_updated_counter // update_step
                         or tree.now_gap <= gap or (use_max_steps and tree.counter >= max_steps)
                         or time.time() - start_time > time_limit)
            add_update(print_message=(verbosity >= 1) if important else (verbosity >= 3))
            if root_only:
                break
        if root_only:
            break

    o4 = problem.objective_mse(solution["X"])
    solution["MSE_in"], solution["MSE_out"], solution["MSE_all"] = o4[1], o4[2], o4[3]
    end_time = time.time()
    instance["run_details"] = dict(
        k=k, m=m, n=n, A=A, indices=indices, num_indices=int(indices.sum()), γ=gamma, node_selection=node_selection,
        bestfirst_depthfirst_cutoff=bestfirst_depthfirst_cutoff, optimality_gap=gap, root_only=root_only,
        altmin_flag=altmin_flag, max_altmin_probability=max_altmin_probability,
        min_altmin_probability=min_altmin_probability, altmin_probability_decay_rate=altmin_probability_decay_rate,
        altmin_root_n_iters=altmin_root_n_iters, use_max_steps=use_max_steps, max_steps=max_steps, time_limit=time_limit,
        use_disjunctive_cuts=use_disjunctive_cuts, disjunctive_cuts_type=disjunctive_cuts_type,
        disjunctive_cuts_breakpoints=disjunctive_cuts_breakpoints, add_Shor_valid_inequalities=add_Shor_valid_inequalities,
        add_Shor_valid_inequalities_fraction=add_Shor_valid_inequalities_fraction,
        add_Shor_valid_inequalities_iterative=add_Shor_valid_inequalities_iterative,
        max_update_Shor_indices_probability=max_update_Shor_indices_probability,
        min_update_Shor_indices_probability=min_update_Shor_indices_probability,
        update_Shor_indices_probability_decay_rate=update_Shor_indices_probability_decay_rate,
        update_Shor_indices_n_minors=update_Shor_indices_n_minors,
        Shor_valid_inequalities_noisy_rank1_num_entries_present=list(Shor_valid_inequalities_noisy_rank1_num_entries_present),
        start_time=start_time, end_time=end_time, time_taken=end_time - start_time,
        solve_time_altmin=solve_time_altmin, dict_solve_times_altmin=dict_solve_times_altmin,
        dict_num_iterations_altmin=dict_num_iterations_altmin, solve_time_relaxation_feasibility=0.0,
        solve_time_relaxation=solve_time_relaxation, dict_solve_times_relaxation=dict_solve_times_relaxation,
        root_node_timeout=root_node_timeout, nodes_explored=tree.nodes_explored, nodes_total=tree.counter,
        warm_start_pool_exhausted=pool_exhausted, **cnt)
    instance["tree"] = tree
    instance["open_nodes"] = [tree.nodes[i] for i in sorted(tree.nodes)]
    instance["problem"] = problem
    if add_Shor_valid_inequalities:
        instance["Shor_info"] = instance_shor                      # BBNodeShorInfo of the root (shared by every node), 0-based
    return solution, printlist, instance
