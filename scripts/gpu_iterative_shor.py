"""GPU diagnostic: branch-and-bound with and without the iterative Shor mode on four small instances (nodes explored, splits,
Shor index updates, final gap); time-boxed at 120 s per run."""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import omc_b200 as omc
from oracle.datagen import generate_matrix_completion_data
omc.init(0)
for (k, n, m, nidx, seed) in ((1, 6, 7, 26, 1), (1, 8, 9, 30, 2), (1, 10, 10, 40, 0), (2, 6, 7, 30, 1)):
    A, mask = generate_matrix_completion_data(k, n, m, nidx, seed)
    kw = dict(node_selection="bestfirst", disjunctive_cuts_type="linear", disjunctive_cuts_breakpoints="smallest_1_eigvec",
              gap=1e-3, max_steps=200, use_max_steps=True, time_limit=120, frontier_batch=4,
              relax_opts=omc.default_opts(eps_abs=1e-7, eps_rel=1e-7, max_iter=20000))
    s0, _, i0 = omc.matrix_completion_branchandbound(k, A, mask, 20.0, **kw)
    s1, _, i1 = omc.matrix_completion_branchandbound(k, A, mask, 20.0, add_Shor_valid_inequalities=True,
                                                     add_Shor_valid_inequalities_iterative=True, update_Shor_indices_n_minors=20, **kw)
    d0, d1 = i0["run_details"], i1["run_details"]
    print((k, n, m, nidx, seed), "plain:", s0["objective"], d0["nodes_explored"], d0["nodes_relax_feasible_split"], i0["tree"].now_gap,
          "| iterative:", s1["objective"], d1["nodes_explored"], d1["nodes_relax_feasible_split"], d1.get("Shor_indices_updates"), i1["tree"].now_gap, flush=True)
