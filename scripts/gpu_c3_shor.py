"""Config 3 as BASELINE.json names it: k = 2, 30 x 30 noisy, gamma = 80, bestfirst, linear2 cuts + Shor valid inequalities
[1, 2, 3, 4].  Root relaxation (with / without the rows), then a time-boxed branch-and-bound with frontier batches."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from oracle.datagen import config_instance

omc.init(0)
k, A, mask, g = config_instance("C3", 0)
n, m = A.shape
p = omc.Problem(k, A, mask, g, "linear2")
t0 = time.time()
minors, soc = omc.shor_constraint_indexes(p, [1, 2, 3, 4])
print(f"C3: {n}x{m} k={k} observed={int(mask.sum())} minors={len(minors)} soc={len(soc)} ({time.time()-t0:.2f}s to enumerate)", flush=True)
EPS = float(os.environ.get("EPS", "1e-6"))
o = omc.default_opts(eps_abs=EPS, eps_rel=EPS, max_iter=int(os.environ.get("MAXIT", "20000")))
r0 = p.relax_batch([[]], o, engine="batched")[0]
print(f"plain root: obj {r0['objective']:.8f} it={r0['iters']} {r0['termination_status']} {r0['solve_time']*1e3:.0f} ms", flush=True)
p.set_shor(minors, soc)
for B in (1, 8):
    fr = p.frontier([[]] * B)
    t0 = time.time(); ms = fr.relax(o); wall = time.time() - t0
    r = fr.fetch()[0]; st = fr.stats()
    print(f"Shor root x{B}: obj {r['objective']:.8f} it={r['iters']} {r['termination_status']} rp={r['res_p']:.2e} rd={r['res_d']:.2e} kernel {ms:.0f} ms wall {wall:.2f}s "
          f"-> {ms/ r['iters'] / B * 1e3:.0f} us/node-iteration, launches {st['launches']}", flush=True)
    fr.close()
if os.environ.get("BNB", "1") == "1":
    t0 = time.time()
    sol, pl, inst = omc.matrix_completion_branchandbound(
        k, A, mask, g, node_selection="bestfirst", disjunctive_cuts_type="linear2", disjunctive_cuts_breakpoints="smallest_1_eigvec",
        add_Shor_valid_inequalities=True, Shor_valid_inequalities_noisy_rank1_num_entries_present=[1, 2, 3, 4],
        time_limit=int(os.environ.get("TL", "240")), frontier_batch=int(os.environ.get("FB", "16")), relax_opts=o, gap=1e-4)
    d = inst["run_details"]
    print(json.dumps(dict(objective=sol["objective"], lower=inst["run_log"][-1][3], upper=inst["run_log"][-1][4], gap=inst["run_log"][-1][5] if len(inst["run_log"][-1]) > 5 else None,
                          nodes_explored=d["nodes_explored"], nodes_total=d["nodes_total"], wall=time.time() - t0,
                          solve_time_relaxation=d["solve_time_relaxation"], root_lb=inst["run_log"][0][3])), flush=True)
