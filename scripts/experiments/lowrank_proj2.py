"""Scratch: cheaper subspace-tracking variants for the minority-side projection inside the oracle ADMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from oracle import relaxation as R

A, mask = bench.c2_instance(0)
cuts = bench.load_frontier_fixture(64)
PMAX = int(os.environ.get("PMAX", 24)); BUF = int(os.environ.get("BUF", 4))
MODE = os.environ.get("MODE", "lobpcg"); PEXP = int(os.environ.get("PEXP", 99))
st = {"b": 0, "Z": [None]*3, "side": [1, 1, -1], "nfull": 0, "nlow": 0, "err": [], "res": [], "spec": []}

def cholqr(B):
    G = B.T @ B
    d = np.sqrt(np.maximum(np.diag(G), 1e-300))
    G = G / np.outer(d, d)
    L = np.linalg.cholesky(G + 1e-14*np.eye(len(G)))
    return np.linalg.solve(L, (B / d).T).T

def track(Vs, Z, b):
    N, p = Z.shape
    W = Vs @ Z
    H = Z.T @ W; H = 0.5*(H+H.T)
    th, G = np.linalg.eigh(H)
    Y = Z @ G; WG = W @ G
    Rr = WG - Y * th
    if MODE == "lobpcg":
        Rr -= Y @ (Y.T @ Rr)
        nr = np.linalg.norm(Rr, axis=0)
        if PEXP < p:   # only the residuals of the top PEXP Ritz pairs
            keep = np.argsort(-th)[:PEXP]
            Rr = Rr[:, keep]; nr = nr[keep]
        keep = nr > 1e-14*np.linalg.norm(Vs)
        Rh = Rr[:, keep] / nr[keep]
        # SVQB
        M = Rh.T @ Rh; ev, Uv = np.linalg.eigh(M); k2 = ev > 1e-10*ev.max()
        Rt = Rh @ (Uv[:, k2] / np.sqrt(ev[k2]))
        Rt -= Y @ (Y.T @ Rt)
        Bs = np.hstack([Y, Rt])
        H2 = Bs.T @ Vs @ Bs; H2 = 0.5*(H2+H2.T)
        th2, G2 = np.linalg.eigh(H2)
        Z = Bs @ G2[:, -p:]; th = th2[-p:]
        Z = cholqr(Z)
        return Z, th, None
    if MODE == "col2x2":
        nr = np.linalg.norm(Rr, axis=0)
        ok = nr > 1e-14*np.linalg.norm(Vs)
        Rh = np.where(ok, Rr / np.where(ok, nr, 1.0), 0.0)
        VR = Vs @ Rh
        gam = np.sum(Rh * VR, axis=0)
        # 2x2 [th, nr; nr, gam]: larger eigenvector
        ang = 0.5*np.arctan2(2*nr, th - gam)
        Z = Y*np.cos(ang) + Rh*np.sin(ang)
        Z = cholqr(Z)
        return Z, None, None
    if MODE == "shift":
        # c from trace: mean of the complement spectrum
        c = -(np.trace(Vs) - th.sum()) / (N - p)
        c = max(c, 1e-12)
        Z = Y + Rr / (th + c)
        Z = cholqr(Z)
        return Z, None, None
    raise SystemExit("mode")

def proj_lowrank(V, b):
    side = st["side"][b]
    Vs = side * V
    Z = st["Z"][b]
    N = V.shape[0]
    if Z is None:
        lam, Q = np.linalg.eigh(Vs)
        st["nfull"] += 1
        r = int((lam > 0).sum())
        if r + BUF <= PMAX:
            p = r + BUF
            st["Z"][b] = Q[:, N - p:]
        Pp = (Q * np.maximum(lam, 0)) @ Q.T
    else:
        st["nlow"] += 1
        p = Z.shape[1]
        Z, th, _ = track(Vs, Z, b)
        W = Vs @ Z; H = Z.T @ W; th, G = np.linalg.eigh(0.5*(H+H.T)); Y = Z @ G
        Rr = W @ G - Y * th
        st["res"].append(np.linalg.norm(Rr[:, th > 0]) / np.linalg.norm(V))
        Pp = (Y * np.maximum(th, 0)) @ Y.T
        r = int((th > 0).sum())
        pn = r + BUF
        if pn > PMAX or pn > p:
            st["Z"][b] = None
        else:
            st["Z"][b] = Y[:, p - pn:]
        if os.environ.get("CHECK"):
            lam, Q = np.linalg.eigh(Vs)
            Pe = (Q * np.maximum(lam, 0)) @ Q.T
            st["err"].append(np.linalg.norm(Pp - Pe) / np.linalg.norm(V))
            if b == 0 and st["nlow"] % 3000 == 0:
                st["spec"].append(lam / np.abs(lam).max())
    return Pp if side > 0 else V + Pp

def psd_project(V):
    b = st["b"] % 3; st["b"] += 1
    V = 0.5*(V+V.T)
    if MODE == "exact":
        lam, Q = np.linalg.eigh(V); return (Q*np.maximum(lam, 0)) @ Q.T
    return proj_lowrank(V, b)
R.psd_project = psd_project
for ni in [int(x) for x in os.environ.get("NODES", "0,5").split(",")]:
    st.update(b=0, Z=[None]*3, nfull=0, nlow=0, err=[], res=[], spec=[])
    r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts[ni], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000))
    print(MODE, "node", ni, "iters", r["iters"], "status", r["status"], "obj %.10f" % r["objective"], "dual %.10f" % r["dual_objective"], "full", st["nfull"], "low", st["nlow"],
          "projerr med/max", (np.median(st["err"]), np.max(st["err"])) if st["err"] else None, "res med/max", (np.median(st["res"]), np.max(st["res"])) if st["res"] else None)
    for s in st["spec"][:2]:
        print("   spectrum blk0 (rel): top", np.round(s[-14:], 6), " bottom", np.round(s[:5], 4), "median", np.median(s))
