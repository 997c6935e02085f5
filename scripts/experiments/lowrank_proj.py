"""Scratch: ADMM with warm-started low-rank (minority-side) PSD projections instead of full eigh."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from oracle import relaxation as R

A, mask = bench.c2_instance(0)
cuts = bench.load_frontier_fixture(64)
PMAX = int(os.environ.get("PMAX", 24)); BUF = int(os.environ.get("BUF", 4)); STEPS = int(os.environ.get("STEPS", 1))
MODE = os.environ.get("MODE", "lobpcg")
st = {"b": 0, "Z": [None]*3, "side": [1, 1, -1], "nfull": 0, "nlow": 0, "err": [], "res": []}

def exact(V):
    lam, Q = np.linalg.eigh(V)
    return lam, Q

def proj_lowrank(V, b):
    """returns P+(V) approx"""
    side = st["side"][b]
    Vs = side * V
    Z = st["Z"][b]
    N = V.shape[0]
    if Z is None:
        lam, Q = exact(Vs)
        st["nfull"] += 1
        r = int((lam > 0).sum())
        if r + BUF <= PMAX:
            p = r + BUF
            st["Z"][b] = Q[:, N - p:]          # top p (includes BUF non-positive ones)
        Pp = (Q * np.maximum(lam, 0)) @ Q.T
    else:
        st["nlow"] += 1
        p = Z.shape[1]
        for s in range(STEPS):
            W = Vs @ Z
            H = Z.T @ W
            th, G = np.linalg.eigh(H)
            Y = Z @ G; Rr = W @ G - Y * th
            # expand
            Rr -= Y @ (Y.T @ Rr)
            Qr, _ = np.linalg.qr(Rr)
            Bs = np.hstack([Y, Qr])
            Bs, _ = np.linalg.qr(Bs)
            H2 = Bs.T @ Vs @ Bs
            th2, G2 = np.linalg.eigh(H2)
            Z = Bs @ G2[:, -p:]
        W = Vs @ Z; H = Z.T @ W; th, G = np.linalg.eigh(0.5*(H+H.T)); Y = Z @ G
        Rr = W @ G - Y * th
        st["res"].append(np.linalg.norm(Rr[:, th > 0]) / np.linalg.norm(V))
        Pp = (Y * np.maximum(th, 0)) @ Y.T
        r = int((th > 0).sum())
        # resize: keep r + BUF
        pn = r + BUF
        if pn > PMAX or pn > p:   # need more vectors: fall back to exact next time
            st["Z"][b] = None
        else:
            st["Z"][b] = Y[:, p - pn:]
        if os.environ.get("CHECK"):
            lam, Q = exact(Vs)
            Pe = (Q * np.maximum(lam, 0)) @ Q.T
            st["err"].append(np.linalg.norm(Pp - Pe) / np.linalg.norm(V))
    return Pp if side > 0 else V + Pp      # P+(V) = V + P+(-V)

def psd_project(V):
    b = st["b"] % 3; st["b"] += 1
    V = 0.5*(V+V.T)
    if MODE == "exact":
        lam, Q = exact(V); return (Q*np.maximum(lam, 0)) @ Q.T
    return proj_lowrank(V, b)
R.psd_project = psd_project
for ni in [0, 5]:
    st.update(b=0, Z=[None]*3, nfull=0, nlow=0, err=[], res=[])
    r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts[ni], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000))
    print(MODE, "node", ni, "iters", r["iters"], "status", r["status"], "obj %.10f" % r["objective"], "dual %.10f" % r["dual_objective"], "full", st["nfull"], "low", st["nlow"],
          "projerr med/max", (np.median(st["err"]), np.max(st["err"])) if st["err"] else None, "res med/max", (np.median(st["res"]), np.max(st["res"])) if st["res"] else None)
