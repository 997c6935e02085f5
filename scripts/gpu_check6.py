"""End-to-end branch-and-bound through the C ABI (Python mirror of the Julia host loop)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from oracle.datagen import config_instance, CONFIGS
omc_b200.init(0)
for cfg, sel, fb, tl in [("C1", "breadthfirst", 1, 60), ("C1", "breadthfirst", 64, 60), ("C1", "bestfirst", 64, 60), ("C2", "bestfirst", 1, 60), ("C2", "bestfirst", 148, 120)]:
    k, A, mask, g = config_instance(cfg, 0)
    t = time.time()
    sol, pl, inst = omc_b200.matrix_completion_branchandbound(
        k, A, mask, g, node_selection=sel, disjunctive_cuts_type=CONFIGS[cfg]["cut_type"],
        disjunctive_cuts_breakpoints="smallest_1_eigvec", time_limit=tl, frontier_batch=fb, use_cutoff=(fb > 1), verbosity=0)
    rd = inst["run_details"]; tr = inst["tree"]
    print(cfg, sel, "batch", fb, f"{time.time()-t:.1f}s", "obj0", sol["objective_initial"], "obj", sol["objective"], "LB", tr.best_lower_bound, "gap", tr.now_gap,
          "explored", rd["nodes_explored"], "total", rd["nodes_total"], "open", len(inst["open_nodes"]),
          {a: b for a, b in rd.items() if a.startswith("nodes_") and a not in ("nodes_explored", "nodes_total")}, "relax_time", round(rd["solve_time_relaxation"], 2), flush=True)
print("DONE")
