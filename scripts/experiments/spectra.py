"""Scratch: spectra of the three PSD-block arguments V along the oracle ADMM trajectory on C2 frontier nodes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from oracle import relaxation as R

A, mask = bench.c2_instance(0)
cuts = bench.load_frontier_fixture(64)
log = []
orig = R.psd_project
state = {"it": 0, "b": 0, "prev": [None]*3}
def psd_project(V):
    lam, Q = np.linalg.eigh(0.5*(V+V.T))
    b = state["b"] % 3; it = state["b"] // 3
    state["b"] += 1
    nrm = np.abs(lam).max()
    npos = int((lam > 0).sum()); nneg = int((lam < 0).sum())
    small = int((np.abs(lam) < 1e-6*nrm).sum()); small4 = int((np.abs(lam) < 1e-4*nrm).sum())
    Qp = Q[:, lam > 0]
    ang = np.nan
    if state["prev"][b] is not None and state["prev"][b].shape[1] == Qp.shape[1] and Qp.shape[1] > 0:
        s = np.linalg.svd(state["prev"][b].T @ Qp, compute_uv=False)
        ang = np.sqrt(max(0.0, 1 - s.min()**2))
    state["prev"][b] = Qp
    if it % 100 == 0 or it < 5:
        log.append((it, b, V.shape[0], npos, nneg, small, small4, ang))
    return (Q*np.maximum(lam, 0)) @ Q.T
R.psd_project = psd_project
for ni in [0, 5, 40]:
    state["b"] = 0; state["prev"] = [None]*3; log.clear()
    r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts[ni], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=5000))
    print("node", ni, "L", len(cuts[ni]), "iters", r["iters"], "status", r["status"], "obj", r["objective"])
    for l in log:
        if l[0] % 500 == 0 or l[0] < 3:
            print("  it %5d blk %d N %3d pos %3d neg %3d |lam|<1e-6: %3d <1e-4: %3d sin(angle) %.2e" % l)
