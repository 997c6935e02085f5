import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["OMP_NUM_THREADS"] = "1"
import numpy as np, bench
from oracle import relaxation as R
from multiprocessing import Pool
A, mask = bench.c2_instance(0)
cuts = bench.load_frontier_fixture(64)
def run(args):
    ni, kw = args
    o = R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=8000, **kw)
    r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts[ni], opts=o)
    return ni, kw, r["iters"], r["status"], r["objective"]
if __name__ == "__main__":
    cfgs = [dict(rho=0.1, adapt_thresh=2.0), dict(rho=0.3, adapt_thresh=2.0), dict(rho=0.1, adapt_thresh=3.0), dict(rho=0.3, adapt_thresh=3.0, adapt_every=50), dict(rho=1.0, adapt_thresh=2.0)]
    nodes = [0, 1, 7, 20, 33, 47, 55, 63]
    with Pool(8) as pool:
        res = pool.map(run, [(ni, c) for c in cfgs for ni in nodes])
    for c in cfgs:
        rs = [r for r in res if r[1] == c]
        print(c, [(r[2], r[3]) for r in rs], "mean", np.mean([r[2] for r in rs]), ["%.6f" % r[4] for r in rs])
