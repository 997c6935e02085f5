"""GPU check of the Shor rows in the batched engine against oracle/shor_relax.py (exact projections)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from oracle import relaxation as R, shor as SI, shor_relax as SR
from oracle.datagen import generate_matrix_completion_data

omc.init(0)

for (k, n, m, nidx, seed, full) in ((1, 5, 6, 18, 3, True), (2, 5, 6, 20, 3, True), (1, 6, 7, 22, 2, False), (2, 8, 9, 40, 1, True), (1, 12, 14, 80, 1, True)):
    A, mask = generate_matrix_completion_data(k, n, m, nidx, seed)
    o = R.Options(eps_abs=1e-7, eps_rel=1e-7, max_iter=40000)
    if full:
        minors = [tuple(v - 1 for v in t) for t in SI.shor_constraint_indexes(mask, [1, 2, 3, 4])]
    else:
        minors = []
    cov = np.zeros((n, m), bool)
    for (i1, i2, j1, j2) in minors:
        cov[i1, j1] = cov[i1, j2] = cov[i2, j1] = cov[i2, j2] = True
    soc = [(i, j) for i in range(n) for j in range(m) if not cov[i, j]]
    t0 = time.time()
    ref = SR.solve_relaxation_shor(A, mask, 20.0, k, minors, soc, opts=o)
    t1 = time.time()
    prob = omc.Problem(k, A, mask, 20.0)
    prob.set_shor(minors, soc)
    fr = prob.frontier([[]])
    opts = omc.default_opts(eps_abs=1e-7, eps_rel=1e-7, max_iter=40000)
    ms = fr.relax(opts)
    r = fr.fetch()[0]
    W, Xt = fr.fetch_shor()
    print(f"k={k} {n}x{m} minors={len(minors)} soc={len(soc)}: oracle {ref['objective']:.9f} it={ref['iters']} ({t1-t0:.1f}s) | gpu {r['objective']:.9f} "
          f"it={r['iters']} st={r['termination_status']} rp={r['res_p']:.2e} rd={r['res_d']:.2e} {ms:.0f} ms | rel={abs(r['objective']-ref['objective'])/abs(ref['objective']):.2e} "
          f"dX={np.abs(r['X']-ref['X']).max():.2e} dW={np.abs(W[0]-ref['W']).max():.2e} dXt={np.abs(Xt[0]-ref['Xt']).max():.2e}", flush=True)
