"""Edge shapes: status / iterations / bound of the tracked and the exact projection path vs each other (and the oracle
for the smallest ones); tracker counters (idle, tracked, full projections) for every shape."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from oracle.datagen import generate_matrix_completion_data
omc.init(0)
shapes = [(2, 2, 1), (2, 3, 2), (3, 7, 2), (5, 5, 1), (4, 9, 4), (3, 3, 1), (4, 4, 1)]
for n, mm, k in itertools.product((6, 8, 10, 12, 16, 24), (1, 2, 3), (1, 2, 3)):
    if k < n:
        shapes.append((n, n * mm, k))
bad = 0
for (n, m, k) in shapes:
    nobs = max(n + m, int(0.6 * n * m))
    A, mask = generate_matrix_completion_data(k, n, m, min(nobs, n * m), 3)
    p = omc.Problem(k, A, mask, 20.0, "linear")
    rs = []
    for ex in (0, 1):
        f = omc.Frontier(p, [[]])
        f.relax(omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=30000, exact_projection=ex))
        r = f.fetch(matrices=False)[0]
        pr = f.profile()[0]
        f.close()
        rs.append((r, pr))
    (rt, pt), (re_, pe) = rs
    ok = rt["termination_status"] == re_["termination_status"] == "OPTIMAL" and abs(rt["objective"] - re_["objective"]) <= 1e-6 * abs(re_["objective"]) \
        and rt["iters"] <= 1.2 * re_["iters"] + 30
    bad += not ok
    print((n, m, k), "OK " if ok else "BAD", "tracked", rt["termination_status"], rt["iters"], "%.10g" % rt["objective"], "idle/lr/full", int(pt[13]), int(pt[14]), int(pt[15]),
          "| exact", re_["termination_status"], re_["iters"], "%.10g" % re_["objective"], flush=True)
    p.close()
print("anomalies:", bad, "of", len(shapes))
