"""Scratch: the exact algorithm planned for the GPU (block LOBPCG step with CholQR2, limited Jacobi, polish)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, bench
from oracle import relaxation as R

A, mask = bench.c2_instance(0)
cuts = bench.load_frontier_fixture(64)
PMAX = int(os.environ.get("PMAX", 16)); BUF = int(os.environ.get("BUF", 2))
SW_C = int(os.environ.get("SW_C", 99)); SW_F = int(os.environ.get("SW_F", 99)); PIV = float(os.environ.get("PIV", 1e-10))
st = {"b": 0, "Z": [None]*3, "side": [1, 1, -1], "nfull": 0, "nlow": 0, "err": [], "steps": 0, "pdist": []}

def rr_pairs(n):
    """round-robin pairs for even n: list of steps, each a list of (p,q)"""
    M = n - 1
    out = []
    for t in range(M):
        prs = [(n - 1, t)]
        for i in range(1, n // 2):
            prs.append(((t + i) % M, (t - i) % M))
        out.append(prs)
    return out
_pairs = {}

def jacobi(Hm, G, idx, sweeps, tol=1e-14):
    """in-place cyclic Jacobi restricted to the index subset idx (even length)"""
    n = len(idx)
    if n < 2: return 0
    if n % 2: raise ValueError
    if n not in _pairs: _pairs[n] = rr_pairs(n)
    fro = np.linalg.norm(Hm)
    done = 0
    for sw in range(sweeps):
        sub = Hm[np.ix_(idx, idx)]
        off = np.sqrt(max(0.0, np.sum(sub*sub) - np.sum(np.diag(sub)**2)))
        if off <= tol * fro: break
        done += 1
        for prs in _pairs[n]:
            st["steps"] += 1
            P = np.array([idx[a] for a, _ in prs]); Q = np.array([idx[b] for _, b in prs])
            apq = Hm[P, Q]; d = Hm[Q, Q] - Hm[P, P]; o = 2*apq
            rr = np.sqrt(d*d + o*o); ok = np.abs(apq) > 1e-300
            rr = np.where(ok, rr, 1.0)
            c2 = np.where(ok, 0.5 + 0.5*np.abs(d)/rr, 1.0); c = np.sqrt(c2)
            s = np.where(ok, np.copysign(0.5*np.abs(o)/rr/c, d*o), 0.0)
            # rows
            rp = Hm[P, :].copy(); rq = Hm[Q, :].copy()
            Hm[P, :] = c[:, None]*rp - s[:, None]*rq; Hm[Q, :] = s[:, None]*rp + c[:, None]*rq
            cp = Hm[:, P].copy(); cq = Hm[:, Q].copy()
            Hm[:, P] = c*cp - s*cq; Hm[:, Q] = s*cp + c*cq
            gp = G[:, P].copy(); gq = G[:, Q].copy()
            G[:, P] = c*gp - s*gq; G[:, Q] = s*gp + c*gq
    return done

def cholqr_drop(Rm, Z):
    """project out Z, scale, Cholesky with pivot threshold; returns orthonormal-ish columns (dropped -> zero) and valid mask"""
    Rm = Rm - Z @ (Z.T @ Rm)
    p = Rm.shape[1]
    M = Rm.T @ Rm
    L = np.zeros((p, p)); valid = np.ones(p, bool)
    scale = max(np.diag(M).max(), 1e-300)
    for j in range(p):
        v = M[j, j] - L[j, :j] @ L[j, :j]
        if v <= PIV * scale or not valid[j]:
            valid[j] = False; L[j, j] = 1.0; L[j+1:, j] = 0.0; L[j, :j] = 0.0
            continue
        L[j, j] = np.sqrt(v)
        L[j+1:, j] = (M[j+1:, j] - L[j+1:, :j] @ L[j, :j]) / L[j, j]
    Rz = np.where(valid, Rm, 0.0)
    Rt = np.linalg.solve(L, Rz.T).T
    Rt = np.where(valid, Rt, 0.0)
    return Rt, valid

ALT = int(os.environ.get("ALT", 0))
_cnt = {"c": 0}
def track(Vs, Z):
    N, p = Z.shape
    W = Vs @ Z
    H = Z.T @ W; H = 0.5*(H + H.T)
    _cnt["c"] += 1
    if ALT and (_cnt["c"] // 3) % ALT != 0:
        th, G = np.linalg.eigh(H)
        order = np.argsort(-th)
        r = int((th > 0).sum())
        if r + 1 > p: return None, None, r
        return Z @ G[:, order], th[order], r
    Rm = W - Z @ H
    nrmV = np.linalg.norm(Vs)
    # first pass: relative pivoting against the largest residual; second pass re-orthonormalises what survived
    Rt, valid = cholqr_drop(Rm, Z)
    global PIV
    piv0 = PIV; PIV = 1e-24
    Rt2, v2 = cholqr_drop(Rt, Z); PIV = piv0
    valid &= v2
    Rt = np.where(valid, Rt2, 0.0)
    WR = Vs @ Rt
    Xc = Z.T @ WR; C = Rt.T @ WR; C = 0.5*(C + C.T)
    n2 = 2*p
    H2 = np.zeros((n2, n2)); H2[:p, :p] = H; H2[:p, p:] = Xc; H2[p:, :p] = Xc.T; H2[p:, p:] = C
    BIG = 1e3 * nrmV
    for j in range(p):
        if not valid[j]:
            H2[p + j, :] = 0; H2[:, p + j] = 0; H2[p + j, p + j] = -BIG
    if SW_F >= 99:
        th, G = np.linalg.eigh(H2)
    else:
        G = np.eye(n2); Hm = H2.copy()
        idxC = [p + j for j in range(p) if valid[j]]
        if len(idxC) % 2: idxC = idxC[:-1] if len(idxC) > 1 else []
        jacobi(Hm, G, idxC, SW_C)
        idxF = list(range(n2)) if n2 % 2 == 0 else list(range(n2))
        jacobi(Hm, G, idxF, SW_F)
        th = np.diag(Hm).copy()
    order = np.argsort(-th)
    r = int((th > 0).sum())
    nvalid = p + int(valid.sum())
    pn = min(r + BUF, nvalid)
    if r + 1 > PMAX: return None, None, r
    pn = min(pn, PMAX)
    sel = order[:pn]
    Bs = np.hstack([Z, Rt])
    return Bs @ G[:, sel], th[sel], r

def proj_lowrank(V, b, exact_now=False):
    side = st["side"][b]
    Vs = side * V
    Z = st["Z"][b]
    N = V.shape[0]
    if Z is None or exact_now:
        lam, Q = np.linalg.eigh(Vs)
        st["nfull"] += 1
        r = int((lam > 0).sum())
        st["Z"][b] = Q[:, N - (r + BUF):][:, ::-1] if r + BUF <= PMAX else None
        Pp = (Q * np.maximum(lam, 0)) @ Q.T
    else:
        st["nlow"] += 1
        Zn, th, r = track(Vs, Z)
        if Zn is None:
            st["Z"][b] = None
            return proj_lowrank(V, b)
        st["pdist"].append(Zn.shape[1])
        Pp = (Zn * np.maximum(th, 0)) @ Zn.T
        st["Z"][b] = Zn
        if os.environ.get("CHECK"):
            lam, Q = np.linalg.eigh(Vs)
            Pe = (Q * np.maximum(lam, 0)) @ Q.T
            st["err"].append(np.linalg.norm(Pp - Pe) / np.linalg.norm(V))
    return Pp if side > 0 else V + Pp

def psd_project(V):
    b = st["b"] % 3; st["b"] += 1
    return proj_lowrank(0.5*(V + V.T), b)
R.psd_project = psd_project
import time
for ni in [int(x) for x in os.environ.get("NODES", "0,5").split(",")]:
    st.update(b=0, Z=[None]*3, nfull=0, nlow=0, err=[], steps=0, pdist=[])
    t0 = time.time()
    r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts[ni], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000))
    print("node", ni, "iters", r["iters"], "status", r["status"], "obj %.10f" % r["objective"], "dual %.10f" % r["dual_objective"], "full", st["nfull"], "low", st["nlow"],
          "projerr med/max", (float(np.median(st["err"])), float(np.max(st["err"]))) if st["err"] else None, "jacobi steps/proj %.1f" % (st["steps"]/max(1, st["nlow"])),
          "orth", [float(np.abs(z.T@z - np.eye(z.shape[1])).max()) if z is not None else None for z in st["Z"]], "p mean %.1f" % np.mean(st["pdist"]), "%.0fs" % (time.time()-t0))
