"""Scratch: does a rank-k primal start (from the zero-filled SVD) avoid the high-rank early ADMM iterations?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from oracle import relaxation as R
from oracle.datagen import config_instance
cfg = os.environ.get("CFG", "C2")
k, A, mask, g = config_instance(cfg, 0)
n, m = A.shape
cuts = bench.load_frontier_fixture(8)[0] if cfg == "C2" and os.environ.get("NODE") else []
cnt = [0]; log = []
def pp(V):
    lam, Q = np.linalg.eigh(0.5*(V+V.T))
    b = cnt[0] % 3; it = cnt[0] // 3; cnt[0] += 1
    if b == 0: log.append((it, int((lam > 0).sum())))
    return (Q*np.maximum(lam, 0)) @ Q.T
R.psd_project = pp
def start_state(c):
    st = R.RelaxState(c)
    Uf, sv, Vt = np.linalg.svd(np.where(mask, A, 0.0) * (A.size / mask.sum()), full_matrices=False)
    U0 = Uf[:, :k]
    X0 = (U0 * sv[:k]) @ Vt[:k]
    Y0 = U0 @ U0.T
    a, sa = c.a, c.sa
    st.X = X0; st.Y = a * Y0; st.U = sa * np.clip(U0, -1, 1)
    st.T = (X0.T @ X0) / a          # Theta~ = Theta / a with Theta = X' Y^+ X = X'X for a projector Y
    st.s1 = np.block([[st.Y, st.X], [st.X.T, st.T]])
    st.s2 = np.block([[st.Y, st.U], [st.U.T, np.eye(k)]])
    st.s3 = c.I3 - st.Y
    st.s4 = max(c.ktr - np.trace(st.Y), 0.0)
    st.s5 = np.clip(st.U, c.lo, c.hi)
    return st
for mode in ("zero", "primal"):
    cnt[0] = 0; log.clear()
    o = R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000)
    c = R.Consts(A, mask, g, k, "linear" if cfg != "C4" else "linear3", cuts, o)
    st = start_state(c) if mode == "primal" else None
    t0 = time.time()
    r = R.solve_relaxation(A, mask, g, k, "linear" if cfg != "C4" else "linear3", cuts, opts=o, state=st)
    big = [it for it, rr in log if rr > 14]
    print(mode, "iters", r["iters"], "st", r["status"], "obj %.8f" % r["objective"], "iterations with r>14:", len(big), "last", (big[-1] if big else None), "r first 12:", [rr for _, rr in log[:12]], "%.0fs" % (time.time() - t0))
