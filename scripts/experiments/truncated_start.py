"""Scratch: early ADMM iterations with the PSD projection truncated to its RCAP largest eigenvalues (what a tracker with
a fixed panel would deliver while the minority side is still wider than the panel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from oracle import relaxation as R
from oracle.datagen import config_instance
cfg = os.environ.get("CFG", "C2"); RCAP = int(os.environ.get("RCAP", 14))
k, A, mask, g = config_instance(cfg, 0)
ctype = {"C1": "linear", "C2": "linear", "C3": "linear2", "C4": "linear3"}[cfg]
cutsets = [[]] + ([bench.load_frontier_fixture(8)[i] for i in (0, 3)] if cfg == "C2" else [])
cnt = [0]; trunc = [0]
def make(mode):
    def pp(V):
        lam, Q = np.linalg.eigh(0.5*(V+V.T))
        b = cnt[0] % 3; it = cnt[0] // 3; cnt[0] += 1
        if mode == "trunc" and it < 150:
            if b < 2:      # natural minority side: positive.  keep only the RCAP largest positive eigenvalues
                if int((lam > 0).sum()) > RCAP:
                    lam2 = lam.copy(); lam2[:len(lam) - RCAP] = np.minimum(lam2[:len(lam) - RCAP], 0.0); trunc[0] += 1
                    return (Q*np.maximum(lam2, 0)) @ Q.T
            else:          # natural minority side: negative.  P+(V) = V - P-(V), P- truncated to the RCAP most negative
                if int((lam < 0).sum()) > RCAP:
                    lam2 = np.where(np.arange(len(lam)) < RCAP, lam, 0.0); trunc[0] += 1
                    return 0.5*(V+V.T) - (Q*np.minimum(lam2, 0)) @ Q.T
        return (Q*np.maximum(lam, 0)) @ Q.T
    return pp
for cuts in cutsets:
    for mode in ("exact", "trunc"):
        cnt[0] = 0; trunc[0] = 0
        R.psd_project = make(mode)
        r = R.solve_relaxation(A, mask, g, k, ctype, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=8000))
        print(cfg, "cuts", len(cuts), mode, "iters", r["iters"], "st", r["status"], "obj %.9f" % r["objective"], "truncated projections", trunc[0], flush=True)

# Outcome (round 1): with EXACT eigenvectors the truncation is harmless (same iteration counts on C1-C4, above).  Driven by
# the kernel's tracker (one LOBPCG step per iteration, panel 16) it was only a 3 % gain on config 2 and it broke config 4
# (minority side up to 50-100 wide for ~80 iterations: the tracked 16 columns do not follow it and the ADMM is led astray;
# root no longer converges in 3000 iterations).  The kernel therefore keeps the full solver for the start phase.
