"""Short batched-engine run for ncu: B config-4 depth-1 nodes, a fixed number of lockstep iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from omc_b200.synthetic import generate_matrix_completion_data
omc.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
import os
if os.environ.get("CFG", "C4") == "C5":
    k, n, m, nidx, ct, labs = 5, 1000, 1000, 200000, "linear", ["left", "right"]
else:
    k, n, m, nidx, ct, labs = 3, 100, 100, 3000, "linear3", ["left", "inner_left", "inner_right", "right"]
A, mask = generate_matrix_completion_data(k, n, m, nidx, 0)
p = omc.Problem(k, A, mask, 80.0, ct)
rng = np.random.default_rng(0)
node_cuts = []
for b in range(B):
    cuts = []
    for l in range(1 + b % 3):
        x = rng.standard_normal(n); x /= np.linalg.norm(x)
        Uh = 0.3 * rng.standard_normal((n, k))
        cuts.append(omc.Cut(p.add_cut(x, Uh), x, Uh, [labs[rng.integers(max(1, len(labs) - 1))] for _ in range(k)]))
    node_cuts.append(cuts)
f = p.frontier(node_cuts, engine="batched")
ms = f.relax(omc.default_opts(eps_abs=1e-12, eps_rel=1e-12, max_iter=iters))
st = f.stats()
print(f"B={B} iters={iters}: {ms:.1f} ms, {ms * 1e3 / iters:.0f} us per lockstep iteration, {ms * 1e6 / st['node_iterations']:.0f} ns per node-iteration, launches {st['launches']}")
f.close(); p.close()
