"""How sensitive is the number of ADMM iterations a node needs (to be certified) to rounding-level perturbations?
The NumPy restatement of the batched engine (oracle/bigblock.py) on config-4 frontier nodes, once as is and once with the data
scaled by (1 + 1e-15): every intermediate result changes in its last bits, nothing else.  Motivation:
scripts/experiments/dmma_panel_kernels_r02.md (a different summation order moved one C5 node from <= 775 to ~3 000 iterations)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import bigblock as Bg
from oracle.cuts import LABELS
from oracle.datagen import generate_matrix_completion_data

root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fx = json.load(open(os.path.join(root, "tests", "golden", "c4_frontier_pool.json")))
A, mask = generate_matrix_completion_data(fx["k"], fx["n"], fx["m"], fx["n_indices"], fx["seed"])
lab = LABELS[fx["cut_type"]]
ids = [i for i, g in enumerate(fx["gpu_first"]) if g["status"] == 0][: int(sys.argv[1]) if len(sys.argv) > 1 else 6]
for q in ids:
    cuts = [(np.array(fx["pool"][e[0]][0]), np.array(fx["pool"][e[0]][1]), [lab[d] for d in e[1:]]) for e in fx["nodes"][q]]
    out = []
    for scale in (1.0, 1.0 + 1e-15, 1.0 - 1e-15):
        t0 = time.time()
        r = Bg.solve_relaxation_big(A * scale, mask, fx["gamma"], fx["k"], fx["cut_type"], cuts,
                                    opts=Bg.BigOptions(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000, cutoff=fx["incumbent"], infeasible_by_bound=True))
        out.append((r["iters"], r["status"], round(r["objective"], 6), round(time.time() - t0, 1)))
    print("node", q, "gpu iters", fx["gpu_first"][q]["iters"], "| restatement (iters, status, objective, s):", out, flush=True)
