"""CPU restatement of the warm-started minority-side PSD projection.  TEST INFRASTRUCTURE ONLY.

The reference projects nothing itself: it hands the whole node program to Mosek (OMC.jl:1859).  The CUDA engine
solves the same program by a conic ADMM whose per-iteration cost is the three PSD-cone projections of
OMC.jl:1554-1556; ``csrc/omc_lowrank.cuh`` replaces the full eigendecomposition of each projection argument V by
one block-LOBPCG step on a tracked basis of the smaller spectral side.  This file states that scheme in NumPy
(same steps, same constants) so that CPU tests can pin its two contracts without a GPU:

  * ``TrackedProjector.project`` stays within ~1e-8 ||V|| of the exact projection along an ADMM trajectory, and
  * the ADMM driven by it reaches the same bound as the ADMM with exact projections (tests/test_oracle.py).

Scheme, for side * V with an orthonormal basis Z (N x p, p = r + BUF <= PM):
    W = V Z;  H = Z'W;  R = W - Z H;  R -= Z (Z'R)                       (block residual, re-orthogonalised)
    R~ = CholQR(R) with a rank guard (pivot < 1e-10 * max diag dropped), a second pass if ill-conditioned
    Rayleigh-Ritz on [Z R~]: two cyclic Jacobi sweeps on the (p + p') x (p + p') matrix
    keep the r + BUF largest Ritz pairs;  P+(side V) ~= sum_{theta_a > 0} theta_a z_a z_a'
A full eigendecomposition (re)starts the basis whenever r + BUF > PM or no guard vector is left.
"""
import numpy as np

BUF = 2          # OMC_LR_BUF
SWEEPS = 2       # OMC_LR_SWEEPS


def _round_robin(n):
    M = n - 1
    steps = []
    for t in range(M):
        prs = [(n - 1, t)] + [((t + i) % M, (t - i) % M) for i in range(1, n // 2)]
        steps.append((np.array([a for a, _ in prs]), np.array([b for _, b in prs])))
    return steps


_RR = {}


def jacobi_sweeps(Hm, sweeps=SWEEPS):
    """Cyclic two-sided Jacobi (round-robin ordering of csrc/omc_device.cuh:jacobi_pair).  Returns (diag, G)."""
    n = Hm.shape[0]
    Hm = Hm.copy(); G = np.eye(n)
    if n % 2:
        raise ValueError("even size expected")
    if n not in _RR:
        _RR[n] = _round_robin(n)
    for _ in range(sweeps):
        for P, Q in _RR[n]:
            apq = Hm[P, Q]; d = Hm[Q, Q] - Hm[P, P]; o = 2 * apq
            ok = np.abs(apq) > 1e-300
            rr = np.where(ok, np.sqrt(d * d + o * o), 1.0)
            c2 = np.where(ok, 0.5 + 0.5 * np.abs(d) / rr, 1.0); c = np.sqrt(c2)
            s = np.where(ok, np.copysign(0.5 * np.abs(o) / rr / c, d * o), 0.0)
            rp = Hm[P, :].copy(); rq = Hm[Q, :].copy()
            Hm[P, :] = c[:, None] * rp - s[:, None] * rq; Hm[Q, :] = s[:, None] * rp + c[:, None] * rq
            cp = Hm[:, P].copy(); cq = Hm[:, Q].copy()
            Hm[:, P] = c * cp - s * cq; Hm[:, Q] = s * cp + c * cq
            gp = G[:, P].copy(); gq = G[:, Q].copy()
            G[:, P] = c * gp - s * gq; G[:, Q] = s * gp + c * gq
    return np.diag(Hm).copy(), G


def cholqr_guarded(Rm, piv_rel, valid_in=None):
    """Cholesky-QR with a rank guard: columns whose pivot <= piv_rel * max diag are dropped (zeroed)."""
    p = Rm.shape[1]
    M = Rm.T @ Rm
    L = np.zeros((p, p)); valid = np.ones(p, bool) if valid_in is None else valid_in.copy()
    dmax = max(np.diag(M).max(), 0.0)
    pmin = dmax
    for j in range(p):
        v = M[j, j] - L[j, :j] @ L[j, :j]
        if not (valid[j] and v > piv_rel * dmax and v > 0):
            valid[j] = False; L[j, j] = 1.0
            continue
        pmin = min(pmin, v)
        L[j, j] = np.sqrt(v)
        L[j + 1:, j] = (M[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    L[:, ~valid] = 0.0; L[~valid, ~valid] = 1.0
    Rt = np.linalg.solve(L, np.where(valid, Rm, 0.0).T).T
    return np.where(valid, Rt, 0.0), valid, pmin < 1e-5 * dmax


def lowrank_step(Vs, Z, pm):
    """One tracking step on Vs = side * V.  Returns (Z_new, theta_new, r, need_full)."""
    N, p = Z.shape
    W = Vs @ Z
    H = Z.T @ W
    R = W - Z @ H
    R = R - Z @ (Z.T @ R)
    Rt, valid, ill = cholqr_guarded(R, 1e-10)
    if ill:
        Rt = Rt - Z @ (Z.T @ Rt)
        Rt, valid, _ = cholqr_guarded(Rt, 1e-24, valid)
    WR = Vs @ Rt
    X = Z.T @ WR; C = Rt.T @ WR
    live = np.concatenate([np.arange(p), p + np.nonzero(valid)[0]])
    nl = len(live); n2 = nl + (nl & 1)
    H2f = np.block([[0.5 * (H + H.T), X], [X.T, 0.5 * (C + C.T)]])
    big = 1e3 * np.linalg.norm(Vs) + 1.0
    H2 = -big * np.eye(n2)
    H2[:nl, :nl] = H2f[np.ix_(live, live)]
    th, G = jacobi_sweeps(H2)
    th = th[:nl]
    order = np.argsort(-th, kind="stable")
    r = int((th > 0).sum())
    need_full = (r + 1 > pm) or (r >= nl) or (2 * (r + BUF) > Vs.shape[0])
    pn = min(r + BUF, nl, pm)
    sel = order[:pn]
    B = np.hstack([Z, Rt])[:, live]
    return B @ G[:nl, sel], th[sel], min(r, pn), need_full


class TrackedProjector:
    """Projection onto the PSD cone of a slowly varying sequence of symmetric matrices (one PSD block)."""

    def __init__(self, pm=16):
        self.pm = pm; self.Z = None; self.side = 1
        self.n_full = 0; self.n_lr = 0

    def _full(self, V):
        lam, Q = np.linalg.eigh(V)
        self.n_full += 1
        npos, nneg = int((lam > 0).sum()), int((lam < 0).sum())
        side = 1 if npos <= nneg else -1
        r = npos if side > 0 else nneg
        N = V.shape[0]
        if r + BUF <= self.pm and 2 * (r + BUF) <= N:   # [Z R~] has 2 p directions: small blocks stay exact
            order = np.argsort(-side * lam, kind="stable")
            self.Z = Q[:, order[:r + BUF]]; self.side = side
        else:
            self.Z = None
        return (Q * np.maximum(lam, 0)) @ Q.T

    def project(self, V, exact=False):
        V = 0.5 * (V + V.T)
        if self.Z is None or exact:
            return self._full(V)
        Zn, th, r, need_full = lowrank_step(self.side * V, self.Z, self.pm)
        if need_full:
            return self._full(V)
        self.n_lr += 1
        self.Z = Zn
        Pp = (Zn[:, :r] * th[:r]) @ Zn[:, :r].T
        return Pp if self.side > 0 else V + Pp
