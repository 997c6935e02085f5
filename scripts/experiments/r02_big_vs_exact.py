"""Scratch (round 2): oracle/bigblock.py (tracked, truncated panel) against oracle/relaxation.py (exact eigh) on config shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import relaxation as R, bigblock as Bg
from oracle.datagen import config_instance
cfg = sys.argv[1]; pm = int(sys.argv[2]) if len(sys.argv) > 2 else 16
k, A, mask, g = config_instance(cfg, 0)
kw = dict(eps_abs=1e-8, eps_rel=1e-8, max_iter=8000)
t0 = time.time(); log = []
rb = Bg.solve_relaxation_big(A, mask, g, k, opts=Bg.BigOptions(pm=pm, verbose=bool(os.environ.get("V")), **kw), log=log)
t1 = time.time()
print(cfg, "big  : iters", rb["iters"], "status", rb["status"], "obj %.10f" % rb["objective"], "lb %.10f" % rb["lower_bound"], "steps", rb["tracker_steps"], "%.1fs" % (t1 - t0), flush=True)
if not os.environ.get("NOEXACT"):
    re_ = R.solve_relaxation(A, mask, g, k, opts=R.Options(**kw))
    print(cfg, "exact: iters", re_["iters"], "status", re_["status"], "obj %.10f" % re_["objective"], "%.1fs" % (time.time() - t1), "rel diff %.2e" % (abs(rb["objective"] - re_["objective"]) / abs(re_["objective"])))
for l in log[:8] + log[8:200:16]:
    print("  it %4d r %3d %3d %3d res %.1e %.1e %.1e" % l)
