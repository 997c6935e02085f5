import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from oracle import relaxation as R, lowrank as LR
A, mask = bench.c2_instance(0)
cuts = bench.load_frontier_fixture(8)
orig = LR.lowrank_step
cnt = [0]
def step(Vs, Z, pm):
    cnt[0] += 1
    if cnt[0] % 1500 in (0, 1, 2):
        W = Vs @ Z; H = Z.T @ W; Rm = W - Z @ H; Rm -= Z @ (Z.T @ Rm)
        nr = np.linalg.norm(Rm, axis=0); th = np.diag(H)
        o = np.argsort(-th)
        print("call", cnt[0], "N", Vs.shape[0], "p", Z.shape[1], "theta/|V|", np.array2string(th[o]/np.linalg.norm(Vs), precision=4, max_line_width=200))
        print("     rel resid norms", np.array2string(nr[o]/nr.max(), precision=3, max_line_width=200), " max/|V| %.2e" % (nr.max()/np.linalg.norm(Vs)))
    return orig(Vs, Z, pm)
LR.lowrank_step = step
r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts[0], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=3100, projection="tracked"))
print(r["iters"], r["status"])
