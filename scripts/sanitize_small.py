"""Small pass over every kernel for compute-sanitizer memcheck (short iteration counts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc, bench
from oracle.datagen import config_instance
omc.init(0)
for cfg, ct, mi in (("C1", "linear", 120), ("C3", "linear2", 60), ("C2", "linear", 45)):
    k, A, mask, g = config_instance(cfg, 0)
    p = omc.Problem(k, A, mask, g, ct, state_pool_capacity=2)
    nodes = [[]]
    if cfg == "C2":
        cuts = bench.load_frontier_fixture(2)
        nodes = [[omc.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
    r = p.relax_batch(nodes, omc.default_opts(max_iter=mi), save_ids=list(range(len(nodes))))
    r2 = p.relax_batch(nodes, omc.default_opts(max_iter=10), warm_ids=list(range(len(nodes))))
    r3 = p.relax_batch(nodes[:1], omc.default_opts(max_iter=8, exact_projection=1))
    lam, vec, bp, feas = omc.smallest_eigvecs_batch(np.stack([q["Y"] for q in r]), np.stack([q["U"] for q in r]), 1)
    U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
    am = omc.alternating_minimization_batch(p, [U0, -U0], max_iters=2)
    t, s = omc.shor_constraint_indexes(p, [4, 3, 2, 1] if cfg != "C2" else [4])
    o = p.objective_mse(np.asfortranarray(r[0]["X"]))
    print(cfg, "ok", r[0]["objective"], r2[0]["iters"], r3[0]["iters"], am[0]["n_iters"], len(t), len(s), o[0], flush=True)
    p.close()
print("DONE")
