import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from oracle.datagen import generate_matrix_completion_data
omc_b200.init(0)
for nidx in [1250, 625, 400]:
    for seed in range(4):
        A, mask = generate_matrix_completion_data(1, 50, 50, nidx, seed)
        t = time.time()
        sol, pl, inst = omc_b200.matrix_completion_branchandbound(1, A, mask, 80.0, node_selection="bestfirst", disjunctive_cuts_type="linear",
            disjunctive_cuts_breakpoints="smallest_1_eigvec", time_limit=25, frontier_batch=148, use_cutoff=True, verbosity=0, stop_at_open_nodes=1200)
        rd = inst["run_details"]; tr = inst["tree"]
        on = inst["open_nodes"]
        print("nidx", nidx, "seed", seed, f"{time.time()-t:.1f}s", "obj0", round(sol["objective_initial"], 5), "obj", round(sol["objective"], 6), "LB", round(tr.best_lower_bound, 6), "gap", f"{tr.now_gap:.2e}",
              "explored", rd["nodes_explored"], "open", len(on), "depths", (min(n.depth for n in on), max(n.depth for n in on)) if on else None, flush=True)
print("DONE")
