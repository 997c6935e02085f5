"""CPU restatement of the batched large-block relaxation engine (csrc/omc_big.cuh).  TEST INFRASTRUCTURE ONLY.
PARITY UNPINNED by the reference (it ships no tests or golden vectors; Julia and Mosek are absent -- oracle/__init__.py): this
restatement is pinned against oracle/relaxation.py (exact eigh projections) and the closed-form roots of oracle/kat.py instead.

Same mathematical program and the same conic ADMM as ``oracle/relaxation.py`` (the reference's node relaxation,
OMC.jl:1491-1499, 1554-1561, 1564-1685, 1848-1856; COSMO/OSQP form, scaling Y~ = a Y, U~ = sqrt(a) U, Theta~ = Theta/a),
written the way the large-block CUDA engine runs it so that sizes beyond an O(N^3) eigendecomposition per iteration
(config 4: N = 200, config 5: N = 2000) can be restated without ``eigh``:

* **v-form.**  Each cone row keeps ``v`` with ``s = P_K(v)`` and ``mu = rho (v - s)``; one ADMM iteration is
  ``v <- v + alpha (z~ - s)`` (identical to ``v = alpha z~ + (1 - alpha) s + mu / rho`` of oracle/relaxation.py).
* **Tracked projections only.**  The minority spectral side of each PSD block argument (positive side of
  ``[Y X; X' Theta]`` and ``[Y U; U' I]``, negative side of ``a I - Y``) is kept as a factor ``Z diag(theta) Z'`` with a
  fixed panel of ``pm`` columns refined by block-LOBPCG steps ``[Z, R~]`` (2 pm x 2 pm Rayleigh-Ritz); while the true
  side is wider than the panel (the first ~100 iterations of a cold start) the projection is the *truncation* to the
  ``pm`` largest Ritz pairs.  ``s`` is never formed: every pass uses ``V`` and the factor.
* Termination, certified bound, cut-off and infeasibility-by-bound are those of oracle/relaxation.py; an OPTIMAL /
  CUTOFF / INFEASIBLE decision is taken only when the trackers are converged (Ritz residual) and hold a guard column
  (fewer than ``pm`` Ritz values on the minority side), otherwise tracker steps are repeated first.

Pinned against ``oracle/relaxation.py`` (exact ``eigh`` projections) at the config 1-4 shapes and against the closed
form ``oracle/kat.py:root_bound_full`` at sizes ``eigh`` cannot reach (tests/test_oracle_big.py).
"""
import numpy as np
from .relaxation import Consts, Options, STATUS_OPTIMAL, STATUS_ITERATION_LIMIT, STATUS_INFEASIBLE
from .objective import compute_SDP_relaxation_objective

STATUS_CUTOFF = 4


class BigOptions(Options):
    def __init__(self, pm=32, pm3=16, window=2, steps_max=3, steps_start=6, track_tol=1e-3, confirm_tol=1e-9, cutoff=np.inf, seed=1, **kw):
        super().__init__(**kw)
        self.pm, self.pm3, self.window, self.steps_max, self.steps_start = pm, pm3, window, steps_max, steps_start
        self.track_tol, self.confirm_tol, self.cutoff, self.seed = track_tol, confirm_tol, cutoff, seed


def hash_unit(i, j, seed):
    """Deterministic pseudo-random value in [-0.5, 0.5) from integers (same integer arithmetic in csrc/omc_big.cuh)."""
    h = (np.uint64(i) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(j) * np.uint64(0xC2B2AE3D27D4EB4F)
         + np.uint64(seed) * np.uint64(0x165667B19E3779F9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(29)
    h = (h * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(32)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


def hash_panel(N, p, seed):
    with np.errstate(over="ignore"):
        i, j = np.meshgrid(np.arange(N, dtype=np.uint64), np.arange(p, dtype=np.uint64), indexing="ij")
        return hash_unit(i, j, seed)


def orthonormalize(Z):
    """CholQR twice (what the engine does: Gram, Cholesky, triangular solve; no Householder)."""
    Q, _ = cholqr(Z, piv_rel=1e-14)
    Q, _ = cholqr(Q, piv_rel=1e-14)
    return Q


def start_basis(N, p, seed):
    return orthonormalize(hash_panel(N, p, seed))


def cholqr(R, piv_rel=1e-12, floor2=0.0, rel_small=0.0, probe=-1):
    """Cholesky-QR with a rank guard on the column-equilibrated Gram matrix.  A column is dropped (zero column) when its
    squared norm is <= floor2, or <= rel_small * (largest squared norm among the columns other than ``probe``), or when its
    pivot in D^-1 R'R D^-1 falls below piv_rel (nearly dependent on earlier columns).  The ``probe`` column is exempt from
    the rel_small rule."""
    p = R.shape[1]
    M = R.T @ R
    d = np.diag(M).copy()
    dres = d.copy()
    if probe >= 0:
        dres[probe] = 0.0
    dmax = max(float(dres.max()), 0.0) if p else 0.0
    colok = (d > floor2) & (d > 0.0) & ((d > rel_small * dmax) | (np.arange(p) == probe))
    sc = np.where(colok, 1.0 / np.sqrt(np.where(colok, d, 1.0)), 0.0)
    M = M * sc[:, None] * sc[None, :]
    L = np.zeros((p, p)); valid = np.ones(p, bool)
    for j in range(p):
        v = M[j, j] - L[j, :j] @ L[j, :j]
        if not (colok[j] and v > piv_rel):
            valid[j] = False
            L[j, j] = 1.0
            continue
        L[j, j] = np.sqrt(v)
        L[j + 1:, j] = (M[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    L[:, ~valid] = 0.0
    L[~valid, :] = 0.0
    L[~valid, ~valid] = 1.0
    Rt = np.linalg.solve(L, (np.where(valid, R, 0.0) * sc).T).T
    return np.where(valid, Rt, 0.0), valid


class Tracker:
    """Top-``p`` eigenpairs of side * V for a slowly varying symmetric V (one PSD block)."""

    def __init__(self, N, pm, side, seed, Z0=None, th0=None, window=0):
        self.N, self.side = N, side
        self.window = window      # > 0: only the r + window leading Ritz pairs (and the probe) take part in a step
        self.p = p = min(pm, N)          # N <= pm: the panel is a complete eigenbasis and the projection is exact
        self.Z = start_basis(N, p, seed) if Z0 is None else Z0.copy()
        self.th = np.zeros(p) if th0 is None else th0.copy()
        self.res = np.inf         # ||V Z - Z diag(theta)||_F of the last step (relative to ||theta||)
        self.nprod = 0

    def step(self, V, tag=0):
        """One block-LOBPCG step on side * V.  Returns the relative Ritz residual before the step.  ``tag`` seeds the
        pseudo-random probe that replaces the last residual column (it * 64 + step * 4 + block in the engine)."""
        Z, sd = self.Z, self.side
        W = sd * (V @ Z)
        H = Z.T @ W
        R = W - Z @ H
        pc = self.p - 1
        # residual window: only the residuals of the positive Ritz pairs and of the first `window` guard columns (and the
        # probe) enter the trial space; the Rayleigh-Ritz still runs over ALL of Z, so no Ritz value ever gets worse -- the
        # far guard columns are simply not refined by their own residuals (they are rotated with everything else)
        p = self.p
        na = p
        if self.window > 0 and tag >= 128:
            na = min(p, int((self.th > 0).sum()) + self.window)
        if na < p:
            keep = R[:, pc].copy() if pc >= na else None
            R[:, na:] = 0.0
            if keep is not None:
                R[:, pc] = keep
        if pc > 0:
            # the trial space [Z, R] stays generic: an eigenvector exactly orthogonal to Z and to every residual (the identity
            # corner of [Y U; U' I] at a node without cuts) would otherwise never be found again once it left the panel
            amp = 1e-3 * np.sqrt(float((H * H).sum()) / self.N)
            with np.errstate(over="ignore"):
                R[:, pc] = amp * hash_unit(np.arange(self.N, dtype=np.uint64), np.uint64(tag), 12345)
        # two explicit projections against Z ("twice is enough"): R = W - Z H carries a component (I - Z'Z) H along Z that
        # can dwarf a converged residual column; one projection leaves (I - Z'Z) times it, which CholQR's normalisation then
        # amplifies by ||H|| / ||R_col|| -- measured: orthonormality of Z lost within 20 iterations
        R = R - Z @ (Z.T @ R)
        R = R - Z @ (Z.T @ R)
        cols = self.th > 0
        if pc > 0:
            cols = cols & (np.arange(self.p) != pc)       # (column p - 1 holds the probe)
        resid = float(np.linalg.norm(R[:, cols]))
        # dropped: residual columns below 1e-10 ||H|| (rounding noise), below 1e-5 of the largest residual column, or nearly
        # dependent on earlier columns (equilibrated pivot below 1e-10)
        hs2 = float((H * H).sum(axis=0).max())
        Rt, valid = cholqr(R, piv_rel=1e-10, floor2=1e-20 * hs2, rel_small=1e-10, probe=pc if pc > 0 else -1)
        W2 = sd * (V @ Rt)
        self.nprod += 2
        G = np.zeros((2 * p, 2 * p))
        G[:p, :p] = 0.5 * (H + H.T)
        Xc = Z.T @ W2
        C = Rt.T @ W2
        G[:p, p:] = Xc; G[p:, :p] = Xc.T
        G[p:, p:] = 0.5 * (C + C.T)
        big = 64.0 * (np.abs(G).max() + 1.0)
        for j in np.nonzero(~valid)[0]:         # dropped directions sink to the bottom of the spectrum
            G[p + j, :] = 0.0; G[:, p + j] = 0.0; G[p + j, p + j] = -big
        lam, Q = np.linalg.eigh(G)
        sel = np.argsort(-lam, kind="stable")[:p]
        Qs = Q[:, sel]
        # sign convention: largest-|.| component of each Ritz coefficient vector positive
        sg = np.sign(Qs[np.abs(Qs).argmax(axis=0), np.arange(p)]); sg[sg == 0] = 1.0
        Qs = Qs * sg
        self.Z = Z @ Qs[:p] + Rt @ Qs[p:]
        self.th = lam[sel]
        scale = max(float(np.linalg.norm(self.th)), 1e-300)
        self.res = resid / scale
        return self.res

    def r(self):
        return int((self.th > 0).sum())

    def factor(self):
        """(Zr, theta_r): the minority-side part  side * sum_{theta>0} theta z z'."""
        keep = self.th > 0
        return self.Z[:, keep], self.side * self.th[keep]

    def lowrank(self):
        Zr, th = self.factor()
        return (Zr * th) @ Zr.T

    def psd_part(self, V):
        F = self.lowrank()
        return F if self.side > 0 else V - F


class BigState:
    """Warm-start record: w, the v-form rows and the tracked factors."""

    def __init__(self, c, o):
        n, m, k = c.n, c.m, c.k
        self.X = np.zeros((n, m)); self.Y = np.zeros((n, n)); self.T = np.zeros((m, m)); self.U = np.zeros((n, k))
        self.V1 = np.zeros((n + m, n + m)); self.V2 = c.E2.copy(); self.V3 = c.I3.copy()
        self.v4 = c.ktr; self.v5 = np.zeros((n, k))
        self.vv = np.zeros((0, k)); self.vg = np.zeros(0)
        # panel widths: 32 columns for [Y X; X' Theta] and [Y U; U' I] (deep nodes of config 4 hold up to ~28 positive
        # eigenvalues), 16 for the negative side of aI - Y (a handful)
        self.tr = [Tracker(n + m, o.pm, +1, o.seed, window=o.window), Tracker(n + k, o.pm, +1, o.seed + 1, window=o.window),
                   Tracker(n, min(o.pm, o.pm3), -1, o.seed + 2, window=o.window)]
        # block 2 starts at E2 = diag(0, I_k): its positive side is spanned by the last k unit vectors
        t2 = self.tr[1]
        q = min(k, t2.p)
        E = hash_panel(n + k, t2.p, o.seed + 1)
        E[n:n + q, :] = 0.0; E[:, :q] = 0.0
        E[n:n + q, :q] = np.eye(q)
        t2.Z = orthonormalize(E); t2.th = np.concatenate([np.ones(q), np.zeros(t2.p - q)])
        self.rho = None

    def extended(self, c):
        add = c.L - self.vv.shape[0]
        if add > 0:
            xn = c.x[-add:]
            v = xn @ self.U
            g = c.beta[-add:] + np.sum(c.alpha[-add:] * v, axis=1) - np.einsum("li,ij,lj->l", xn, self.Y, xn)
            self.vv = np.vstack([self.vv, np.clip(v, c.lb[-add:], c.ub[-add:])])
            self.vg = np.concatenate([self.vg, np.maximum(g, 0.0)])
        return self


def solve_relaxation_big(A, mask, gamma, k, cut_type=None, cuts=(), opts=None, state=None, log=None):
    """One node through the large-block engine's algorithm.  Returns the dict of oracle/relaxation.solve_relaxation."""
    o = opts or BigOptions()
    A = np.asarray(A, dtype=float); mask = np.asarray(mask, dtype=bool)
    c = Consts(A, mask, gamma, k, cut_type, cuts, o)
    n, m, L = c.n, c.m, c.L
    Mk = c.Mk
    qX = -Mk * A
    st = (state if state is not None else BigState(c, o)).extended(c)
    rho = st.rho if st.rho is not None else o.rho
    sig, al = o.sigma, o.alpha
    G = c.gram(); r = G.shape[0]
    eyen, eyem = np.eye(n), np.eye(m)
    t1, t2, t3 = st.tr
    status = STATUS_ITERATION_LIMIT
    res_p = res_d = np.inf
    bound = -np.inf
    nsteps_total = 0
    it = 0
    confirm = False
    for it in range(1, o.max_iter + 1):
        # ---- current projections from the factors
        F1 = t1.lowrank(); F2 = t2.lowrank(); F3 = t3.lowrank()
        s1, s2, s3 = F1, F2, st.V3 - F3
        s4 = max(st.v4, 0.0); s5 = np.clip(st.v5, c.lo, c.hi)
        sv = np.clip(st.vv, c.lb, c.ub); sg = np.maximum(st.vg, 0.0)
        # ---- w-update: t/rho = v - 2 s (+ b)
        gX, gY, gT, gU = c.At(rho * (st.V1 - 2 * s1), rho * (st.V2 - 2 * s2 + c.E2), rho * (st.V3 - 2 * s3 + c.I3),
                              rho * (st.v4 - 2 * s4 + c.ktr), rho * (st.v5 - 2 * s5),
                              rho * (st.vv - 2 * sv), rho * (st.vg - 2 * sg + c.beta))
        dYU = sig + 3.0 * rho
        Xt = (sig * st.X - qX + gX) / (Mk + sig + 2.0 * rho)
        Tt = (sig * st.T - c.cT * eyem + gT) / (sig + rho)
        Yt = (sig * st.Y + gY) / dYU
        Ut = (sig * st.U + gU) / dYU
        cw = np.linalg.solve(np.eye(r) * (dYU / rho) + G, c.R(Yt, Ut))
        cY, cU = c.Rt(cw)
        Yt = Yt - cY; Ut = Ut - cU
        z1, z2, z3, z4, z5, zv, zg = c.S(Xt, Yt, Tt, Ut)
        # ---- v <- v + alpha (z~ - s);  w <- alpha w~ + (1 - alpha) w
        st.V1 = st.V1 + al * (z1 - s1); st.V2 = st.V2 + al * (z2 - s2); st.V3 = st.V3 + al * (z3 - s3)
        st.v4 = st.v4 + al * (z4 - s4); st.v5 = st.v5 + al * (z5 - s5)
        st.vv = st.vv + al * (zv - sv); st.vg = st.vg + al * (zg - sg)
        st.X = al * Xt + (1 - al) * st.X; st.Y = al * Yt + (1 - al) * st.Y
        st.T = al * Tt + (1 - al) * st.T; st.U = al * Ut + (1 - al) * st.U
        # ---- tracker steps on the new arguments
        for bidx, (tr, V) in enumerate(((t1, st.V1), (t2, st.V2), (t3, st.V3))):
            ns = o.steps_start if it == 1 else 1
            last = it >= o.max_iter
            conf = confirm or last                  # decision pending: tight tolerance until the next scheduled check,
            check_it = it % o.check_every == 0 or last   # steps_start steps allowed in the iteration of that check
            tol = o.confirm_tol if conf else o.track_tol
            qmax = o.steps_start if (conf and check_it) else o.steps_max
            q = 0
            while True:
                res = tr.step(V, it * 64 + q * 4 + bidx); q += 1
                # res is the residual BEFORE the step: one more step measures the new basis only if asked for
                if q >= ns and (res <= tol or q >= qmax):
                    break
            nsteps_total += q
        if log is not None:
            log.append((it, t1.r(), t2.r(), t3.r(), t1.res, t2.res, t3.res))

        if it % o.check_every == 0 or it == o.max_iter:
            was_confirm = confirm
            confirm = False
            for t in st.tr:               # keep the tracked bases orthonormal over thousands of updates
                t.Z, _ = cholqr(t.Z, piv_rel=1e-14)
            s1 = t1.lowrank(); s2 = t2.lowrank(); s3 = st.V3 - t3.lowrank()
            s4 = max(st.v4, 0.0); s5 = np.clip(st.v5, c.lo, c.hi)
            sv = np.clip(st.vv, c.lb, c.ub); sg = np.maximum(st.vg, 0.0)
            m1 = rho * (st.V1 - s1); m2 = rho * (st.V2 - s2); m3 = rho * (st.V3 - s3)
            m4 = rho * (st.v4 - s4); m5 = rho * (st.v5 - s5); mv = rho * (st.vv - sv); mg = rho * (st.vg - sg)
            w = c.S(st.X, st.Y, st.T, st.U)
            sblk = (s1, s2, s3, np.array([s4]), s5, sv, sg)
            rp = max(np.abs(np.asarray(wi) - si).max() for wi, si in zip(w, sblk) if si.size)
            gX, gY, gT, gU = c.At(m1, m2, m3, m4, m5, mv, mg)
            PX = Mk * st.X
            rd = max(np.abs(PX + qX - gX).max(), np.abs(gY).max(), np.abs(c.cT * eyem - gT).max(), np.abs(gU).max())
            n_p = max(max(np.abs(si).max() for si in sblk if si.size), c.a, c.ktr, np.abs(c.beta).max() if L else 0.0)
            n_d = max(np.abs(PX).max(), np.abs(qX).max(), c.cT, np.abs(gX).max(), np.abs(gY).max(),
                      np.abs(gT).max(), np.abs(gU).max())
            res_p, res_d = rp, rd
            dual = (-0.5 * float(np.sum(Mk * st.X * st.X)) + c.c0 + np.trace(m2[n:, n:]) + c.a * np.trace(m3)
                    + c.ktr * m4 + float(c.beta @ mg) - float(np.sum(np.where(m5 < 0, m5 * c.lo, m5 * c.hi)))
                    - float(np.sum(np.where(mv < 0, mv * c.lb, mv * c.ub))))
            obj_p = 0.5 * float(np.sum(Mk * (st.X - A) ** 2)) + c.cT * float(np.trace(st.T))
            ub = o.cutoff if o.cutoff < 1e299 else 2.0 * max(abs(obj_p), abs(dual)) + 1.0
            w1 = lambda trTb: n * c.ktr + np.sqrt(n * m * c.ktr * trTb) + m * trTb + n * c.k * c.sa
            bound_now = dual - rd * w1(ub / c.cT)
            bound_c0 = dual - rd * w1(c.c0 / c.cT)
            guard_ok = all(t.r() < t.p or t.p == t.N for t in st.tr)
            tracked_ok = (was_confirm or it >= o.max_iter) and guard_ok and max(t.res for t in st.tr) <= 10 * o.confirm_tol
            if tracked_ok:            # mu is in the dual cone only then: the certified bound comes from such checks only
                bound = max(bound, bound_now)
            if o.verbose:
                print(f"it {it:6d} rp {rp:.3e} rd {rd:.3e} rho {rho:.3e} r {[t.r() for t in st.tr]} res {[f'{t.res:.1e}' for t in st.tr]}")
            decision = None
            if rp <= o.eps_abs + o.eps_rel * n_p and rd <= o.eps_abs + o.eps_rel * n_d:
                decision = STATUS_OPTIMAL
            elif o.cutoff < 1e299 and bound_now > o.cutoff:
                decision = STATUS_CUTOFF
            elif o.infeasible_by_bound and L > 0 and bound_c0 > c.c0 * (1.0 + 1e-9) + 1e-12:
                decision = STATUS_INFEASIBLE
            if decision is not None:
                if tracked_ok or it >= o.max_iter:
                    status = decision if tracked_ok else STATUS_ITERATION_LIMIT
                    break
                confirm = True        # the trackers run at confirm_tol until the next scheduled check, which re-decides
                continue
            if o.adaptive_rho and it % o.adapt_every == 0:
                ratio = np.sqrt((rp / max(n_p, 1e-12)) / max(rd / max(n_d, 1e-12), 1e-30))
                if ratio > o.adapt_thresh or ratio < 1.0 / o.adapt_thresh:
                    rho_new = float(np.clip(rho * ratio, 1e-6, 1e6))
                    cfac = rho / rho_new        # mu fixed: v <- s + (rho / rho_new)(v - s)
                    st.V1 = s1 + cfac * (st.V1 - s1); st.V2 = s2 + cfac * (st.V2 - s2); st.V3 = s3 + cfac * (st.V3 - s3)
                    st.v4 = s4 + cfac * (st.v4 - s4); st.v5 = s5 + cfac * (st.v5 - s5)
                    st.vv = sv + cfac * (st.vv - sv); st.vg = sg + cfac * (st.vg - sg)
                    # the eigenvectors are unchanged; minority Ritz values stay, the others scale
                    for t in st.tr:           # (side -1 tracks -V: there the minority values are the mu / rho part)
                        t.th = np.where(t.th > 0, t.th, cfac * t.th) if t.side > 0 else np.where(t.th > 0, cfac * t.th, t.th)
                    rho = rho_new
    st.rho = rho
    X, Y, T, U = st.X, st.Y / c.a, st.T * c.a, st.U / c.sa
    obj = compute_SDP_relaxation_objective(X, Y, T, U, A, mask, gamma)
    return dict(status=status, feasible=status != STATUS_INFEASIBLE, objective=obj, lower_bound=bound,
                X=X, Y=Y, Theta=T, U=U, iters=it, res_p=res_p, res_d=res_d, state=st, rho=rho, consts=c,
                tracker_steps=nsteps_total)
