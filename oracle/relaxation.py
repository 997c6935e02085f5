"""CPU restatement of the per-node convex relaxation.  TEST INFRASTRUCTURE ONLY.

The mathematical program is the reference's (disjunctive path, no Shor rows):

    min  1/2 sum_{(i,j) in I} (A_ij - X_ij)^2 + tr(Theta)/(2 gamma)        OMC.jl:1848-1856
    s.t. [Y X; X' Theta] >= 0                                               OMC.jl:1554
         [Y U; U' I_k]   >= 0                                               OMC.jl:1555
         I - Y           >= 0                                               OMC.jl:1556
         tr Y <= k                                                          OMC.jl:1558
         lo <= U <= 1 (lo = 0 on rows n-k+j.. of column j, else -1)         OMC.jl:1561, 1442-1449
         per cut: lb <= x'U_j <= ub,  sum_j(alpha_j x'U_j + beta_j) >= x'Yx OMC.jl:1564-1685
         ||U_j|| <= 1                                                       OMC.jl:1831-1835

The last row is implied by the second and third (Y >= UU', Y <= I  =>
||U_j||^2 <= lambda_max(UU') <= lambda_max(Y) <= 1), so it never changes the
optimum and is not carried as a separate cone; ``certificate`` reports
``max_j ||U_j|| - 1`` so tests can see it hold.

The reference hands the program to Mosek (closed, absent here).  This file
solves it with the conic ADMM the CUDA engine also runs (COSMO/OSQP form
``min 1/2 w'Pw + q'w  s.t. Aw + s = b, s in K``, written in matrix blocks) after
the change of variables ``Y~ = a Y, U~ = sqrt(a) U, Theta~ = Theta / a`` with
``a = n`` (a congruence of both PSD blocks, so the cones are unchanged; it only
balances the entry magnitudes 1/n, 1, n of Y, X, Theta).  Independent of the
algorithm, ``certificate`` checks a returned point against the KKT conditions
of the program (primal feasibility by eigenvalues, dual cone membership,
stationarity, duality gap) -- the pin that replaces reference golden vectors.

Returned objective: recomputed from the primal point, as OMC.jl:1882-1895.
"""
import numpy as np
from .cuts import cut_rows
from .objective import compute_SDP_relaxation_objective

STATUS_OPTIMAL = 0          # -> MOI.OPTIMAL
STATUS_ITERATION_LIMIT = 1  # -> MOI.SLOW_PROGRESS with values (feasible = true, OMC.jl:1871-1877)
STATUS_INFEASIBLE = 2       # -> MOI.INFEASIBLE
STATUS_TIME_LIMIT = 3       # -> MOI.TIME_LIMIT


def default_U_lower(n, k):
    """OMC.jl:1442-1449 (0-based: rows n-k+j .. n-1 of column j are >= 0)."""
    lo = -np.ones((n, k))
    for j in range(k):
        lo[n - k + j:, j] = 0.0
    return lo


def psd_project(V):
    """Frobenius projection of a symmetric matrix onto the PSD cone."""
    lam, Q = np.linalg.eigh(0.5 * (V + V.T))
    return (Q * np.maximum(lam, 0.0)) @ Q.T


class Options:
    def __init__(self, eps_abs=1e-7, eps_rel=1e-7, max_iter=20000, rho=0.3, sigma=1e-6,
                 alpha=1.6, check_every=25, adapt_every=100, adaptive_rho=True,
                 eps_inf=1e-6, eps_inf_loose=1e-3, infeasible_by_bound=False, fix_linear3_right=False, scale=None, verbose=False, projection="exact", pm=16,
                 adapt_thresh=5.0):
        self.__dict__.update(locals()); del self.__dict__["self"]


class Consts:
    """Constants of the scaled program for one node."""

    def __init__(self, A, mask, gamma, k, cut_type, cuts, o, U_lower=None):
        n, m = A.shape
        self.n, self.m, self.k = n, m, k
        self.a = a = float(n) if o.scale is None else float(o.scale)
        self.sa = sa = np.sqrt(a)
        self.Mk = mask.astype(float)
        self.A = A
        self.cT = a / (2.0 * gamma)
        self.c0 = 0.5 * float(np.sum((A * A)[mask]))
        lo = default_U_lower(n, k) if U_lower is None else U_lower
        self.lo, self.hi = sa * lo, sa * np.ones((n, k))
        rows = cut_rows(cut_type, list(cuts), o.fix_linear3_right) if len(cuts) else None
        self.L = 0 if rows is None else rows["x"].shape[0]
        if rows is not None:
            self.x = rows["x"]; self.lb = sa * rows["lb"]; self.ub = sa * rows["ub"]
            self.alpha = sa * rows["alpha"]; self.beta = a * rows["beta"]
        else:
            self.x = np.zeros((0, n)); self.lb = self.ub = self.alpha = np.zeros((0, k)); self.beta = np.zeros(0)
        self.I3 = a * np.eye(n)
        self.ktr = a * k
        self.E2 = np.zeros((n + k, n + k)); self.E2[n:, n:] = np.eye(k)

    # s = b - A w
    def S(self, X, Y, T, U):
        n, k = self.n, self.k
        s1 = np.block([[Y, X], [X.T, T]])
        s2 = np.block([[Y, U], [U.T, np.eye(k)]])
        s3 = self.I3 - Y
        s4 = self.ktr - np.trace(Y)
        sv = self.x @ U
        sg = self.beta + np.sum(self.alpha * sv, axis=1) - np.einsum("li,ij,lj->l", self.x, Y, self.x)
        return s1, s2, s3, s4, U.copy(), sv, sg

    # A' t   (A = -ds/dw)
    def At(self, t1, t2, t3, t4, t5, tv, tg):
        n = self.n
        gX = -(t1[:n, n:] + t1[n:, :n].T)
        gY = -t1[:n, :n] - t2[:n, :n] + t3 + t4 * np.eye(n) + (self.x.T * tg) @ self.x
        gT = -t1[n:, n:]
        gU = -(t2[:n, n:] + t2[n:, :n].T) - t5 - self.x.T @ (tv + tg[:, None] * self.alpha)
        return gX, gY, gT, gU

    # dense rows (trace row, cut rows) restricted to (Y, U):  R [Y;U]  and  R' c
    def R(self, Y, U):
        xv = self.x @ U
        return np.concatenate([[np.trace(Y)], (-xv).reshape(-1),
                               -np.sum(self.alpha * xv, axis=1) + np.einsum("li,ij,lj->l", self.x, Y, self.x)])

    def Rt(self, c):
        L, k, n = self.L, self.k, self.n
        cv = c[1:1 + L * k].reshape(L, k); cg = c[1 + L * k:]
        return c[0] * np.eye(n) + (self.x.T * cg) @ self.x, -self.x.T @ (cv + cg[:, None] * self.alpha)

    def gram(self):
        r = 1 + self.L * (self.k + 1)
        G = np.zeros((r, r))
        for i in range(r):
            e = np.zeros(r); e[i] = 1.0
            G[:, i] = self.R(*self.Rt(e))
        return G


class RelaxState:
    """(w, s, mu) of the ADMM in matrix blocks (scaled variables) -- also the warm-start record."""

    def __init__(self, c):
        n, m, k, L = c.n, c.m, c.k, 0      # cold start: no cut rows yet; extended() adds them
        self.X = np.zeros((n, m)); self.Y = np.zeros((n, n)); self.T = np.zeros((m, m)); self.U = np.zeros((n, k))
        self.s1 = np.zeros((n + m, n + m)); self.m1 = np.zeros((n + m, n + m))
        self.s2 = c.E2.copy(); self.m2 = np.zeros((n + k, n + k))
        self.s3 = c.I3.copy(); self.m3 = np.zeros((n, n))
        self.s4 = c.ktr; self.m4 = 0.0
        self.s5 = np.zeros((n, k)); self.m5 = np.zeros((n, k))
        self.sv = np.zeros((L, k)); self.mv = np.zeros((L, k))
        self.sg = np.zeros(L); self.mg = np.zeros(L)
        self.rho = None

    def extended(self, c):
        """Child warm start: parent's state plus fresh rows for the cuts the child adds."""
        st = self.copy()
        add = c.L - st.sv.shape[0]
        if add > 0:
            k = c.k
            xn = c.x[-add:]
            v = xn @ st.U
            g = c.beta[-add:] + np.sum(c.alpha[-add:] * v, axis=1) - np.einsum("li,ij,lj->l", xn, st.Y, xn)
            st.sv = np.vstack([st.sv, np.clip(v, c.lb[-add:], c.ub[-add:])]); st.mv = np.vstack([st.mv, np.zeros((add, k))])
            st.sg = np.concatenate([st.sg, np.maximum(g, 0.0)]); st.mg = np.concatenate([st.mg, np.zeros(add)])
        return st

    def copy(self):
        c = RelaxState.__new__(RelaxState)
        for key, v in self.__dict__.items():
            c.__dict__[key] = v.copy() if isinstance(v, np.ndarray) else v
        return c


def solve_relaxation(A, mask, gamma, k, cut_type=None, cuts=(), opts=None, state=None, U_lower=None):
    """Solve one node.  Returns a dict mirroring OMC.jl:1860-1919 plus diagnostics.

    ``cuts``: list of (x, Uhat_or_vhat, dirs) as in BBNodeDisjunctiveCuts.cuts.
    ``state``: a parent's RelaxState for a warm start (None = cold start from zero).
    """
    o = opts or Options()
    A = np.asarray(A, dtype=float); mask = np.asarray(mask, dtype=bool)
    c = Consts(A, mask, gamma, k, cut_type, cuts, o, U_lower)
    n, m, L = c.n, c.m, c.L
    Mk = c.Mk
    qX = -Mk * A
    st = (state if state is not None else RelaxState(c)).extended(c)
    rho = st.rho if st.rho is not None else o.rho
    sig, al = o.sigma, o.alpha
    G = c.gram()
    r = G.shape[0]
    eyen, eyem = np.eye(n), np.eye(m)

    status = STATUS_ITERATION_LIMIT
    res_p = res_d = np.inf
    it = 0
    # projection = "tracked": the warm-started low-rank projection of csrc/omc_lowrank.cuh (oracle/lowrank.py); every
    # termination decision is then re-taken after one iteration with exact projections, as the kernel does
    trackers = None
    if o.projection == "tracked":
        from .lowrank import TrackedProjector
        trackers = [TrackedProjector(o.pm) for _ in range(3)]
    exact_iter = force_check = False
    node_exact = False      # set when a tracked node looks infeasible: exact projections from then on
    n_suspect = 0
    bound_max = -np.inf     # largest certified bound seen at a check (infeasible_by_bound diagnostics)
    for it in range(1, o.max_iter + 1):
        # ---- w-update: (P + sigma I + rho A'A) w~ = sigma w - q + A'(rho (b - s) + mu)
        gX, gY, gT, gU = c.At(st.m1 - rho * st.s1, st.m2 + rho * (c.E2 - st.s2), st.m3 + rho * (c.I3 - st.s3),
                              st.m4 + rho * (c.ktr - st.s4), st.m5 - rho * st.s5,
                              st.mv - rho * st.sv, st.mg + rho * (c.beta - st.sg))
        dYU = sig + 3.0 * rho                                 # Y: PSD1+PSD2+PSD3 ; U: 2 (PSD2) + 1 (box)
        Xt = (sig * st.X - qX + gX) / (Mk + sig + 2.0 * rho)
        Tt = (sig * st.T - c.cT * eyem + gT) / (sig + rho)
        Yt = (sig * st.Y + gY) / dYU
        Ut = (sig * st.U + gU) / dYU
        # Woodbury for the dense rows: (D + rho R'R)^-1 = D^-1 - D^-1 R'(I/rho + R D^-1 R')^-1 R D^-1
        cw = np.linalg.solve(np.eye(r) * (dYU / rho) + G, c.R(Yt, Ut))
        cY, cU = c.Rt(cw)
        Yt = Yt - cY
        Ut = Ut - cU
        # ---- s~ = b - A w~ ; relaxation ; projection ; dual update
        z1, z2, z3, z4, z5, zv, zg = c.S(Xt, Yt, Tt, Ut)
        v1 = al * z1 + (1 - al) * st.s1 + st.m1 / rho
        v2 = al * z2 + (1 - al) * st.s2 + st.m2 / rho
        v3 = al * z3 + (1 - al) * st.s3 + st.m3 / rho
        v4 = al * z4 + (1 - al) * st.s4 + st.m4 / rho
        v5 = al * z5 + (1 - al) * st.s5 + st.m5 / rho
        vv = al * zv + (1 - al) * st.sv + st.mv / rho
        vg = al * zg + (1 - al) * st.sg + st.mg / rho
        st.X = al * Xt + (1 - al) * st.X; st.Y = al * Yt + (1 - al) * st.Y
        st.T = al * Tt + (1 - al) * st.T; st.U = al * Ut + (1 - al) * st.U
        m_old = (st.m1, st.m2, st.m3, st.m4, st.m5, st.mv, st.mg)
        if trackers is None:
            st.s1 = psd_project(v1); st.s2 = psd_project(v2); st.s3 = psd_project(v3)
        else:
            st.s1, st.s2, st.s3 = (tr.project(v, exact=exact_iter or node_exact) for tr, v in zip(trackers, (v1, v2, v3)))
        st.s4 = max(v4, 0.0)
        st.s5 = np.clip(v5, c.lo, c.hi)
        st.sv = np.clip(vv, c.lb, c.ub); st.sg = np.maximum(vg, 0.0)
        st.m1 = rho * (v1 - st.s1); st.m2 = rho * (v2 - st.s2); st.m3 = rho * (v3 - st.s3)
        st.m4 = rho * (v4 - st.s4); st.m5 = rho * (v5 - st.s5)
        st.mv = rho * (vv - st.sv); st.mg = rho * (vg - st.sg)

        if it % o.check_every == 0 or it == o.max_iter or force_check:
            provisional = trackers is not None and not exact_iter and not node_exact
            exact_iter = force_check = False
            w = c.S(st.X, st.Y, st.T, st.U)                                       # b - A w
            sblk = (st.s1, st.s2, st.s3, np.array([st.s4]), st.s5, st.sv, st.sg)
            rp = max(np.abs(np.asarray(wi) - si).max() for wi, si in zip(w, sblk) if si.size)
            gX, gY, gT, gU = c.At(st.m1, st.m2, st.m3, st.m4, st.m5, st.mv, st.mg)
            PX = Mk * st.X
            rd = max(np.abs(PX + qX - gX).max(), np.abs(gY).max(), np.abs(c.cT * eyem - gT).max(), np.abs(gU).max())
            n_p = max(max(np.abs(si).max() for si in sblk if si.size), c.a, c.ktr,
                      np.abs(c.beta).max() if L else 0.0)
            n_d = max(np.abs(PX).max(), np.abs(qX).max(), c.cT, np.abs(gX).max(), np.abs(gY).max(),
                      np.abs(gT).max(), np.abs(gU).max())
            res_p, res_d = rp, rd
            if o.verbose:
                print(f"it {it:6d} rp {rp:.3e} rd {rd:.3e} rho {rho:.3e}")
            if rp <= o.eps_abs + o.eps_rel * n_p and rd <= o.eps_abs + o.eps_rel * n_d:
                if provisional and it < o.max_iter:
                    exact_iter = force_check = True
                    continue
                status = STATUS_OPTIMAL
                break
            # ---- infeasibility by bound (optional; kernel: -DOMC_INFEASIBLE_BY_BOUND): a feasible node has
            # p* <= c0 = 1/2 ||P_Omega(A)||^2 (X = 0, Theta = 0 with any feasible (Y, U)), hence ||w*||_1 <= w1(c0); a
            # certified lower bound  dual - ||r_d||_inf w1(c0)  above c0 contradicts feasibility.  mu must be in the dual
            # cone: under tracked projections the decision is re-taken after one exact iteration, like OPTIMAL.
            if o.infeasible_by_bound and L > 0:
                dual_ = (-0.5 * float(np.sum(Mk * st.X * st.X)) + c.c0
                         + np.trace(st.m2[n:, n:]) + c.a * np.trace(st.m3) + c.ktr * st.m4 + float(c.beta @ st.mg)
                         - float(np.sum(np.where(st.m5 < 0, st.m5 * c.lo, st.m5 * c.hi)))
                         - float(np.sum(np.where(st.mv < 0, st.mv * c.lb, st.mv * c.ub))))
                trTb = c.c0 / c.cT
                w1 = n * c.ktr + np.sqrt(n * m * c.ktr * trTb) + m * trTb + n * c.k * c.sa
                bound_max = max(bound_max, dual_ - rd * w1)
                if dual_ - rd * w1 > c.c0 * (1.0 + 1e-9) + 1e-12:
                    if provisional and it < o.max_iter:
                        exact_iter = force_check = True
                        continue
                    status = STATUS_INFEASIBLE
                    break
            # ---- primal infeasibility (COSMO sec. 5.2): dmu in the polar cone, A'dmu ~ 0, support - b'dmu < 0
            if L > 0:
                d = [np.asarray(a_) - np.asarray(b_) for a_, b_ in
                     zip((st.m1, st.m2, st.m3, st.m4, st.m5, st.mv, st.mg), m_old)]
                nrm = max(np.abs(di).max() for di in d if di.size)
                if nrm > 1e-14:
                    gX, gY, gT, gU = c.At(d[0], d[1], d[2], float(d[3]), d[4], d[5], d[6])
                    atn = max(np.abs(gX).max(), np.abs(gY).max(), np.abs(gT).max(), np.abs(gU).max())
                    sup = (np.sum(np.where(d[4] > 0, c.hi * d[4], c.lo * d[4]))
                           + np.sum(np.where(d[5] > 0, c.ub * d[5], c.lb * d[5])))
                    bdy = (np.trace(d[1][n:, n:]) + c.a * np.trace(d[2]) + c.ktr * float(d[3])
                           + float(c.beta @ d[6]))
                    # tracked projections leave an error floor of ~1e-5 on A'dmu: a node that LOOKS infeasible at a
                    # loose tolerance finishes on exact projections, where the strict certificate can be met
                    if (trackers is not None and not node_exact and atn <= o.eps_inf_loose * nrm
                            and max(float(d[3]), d[6].max()) <= o.eps_inf_loose * nrm and sup - bdy < -o.eps_inf_loose * nrm):
                        node_exact = True
                        n_suspect += 1
                    if atn <= o.eps_inf * nrm:
                        cone_ok = (np.linalg.eigvalsh(d[0]).max() <= o.eps_inf * nrm
                                   and np.linalg.eigvalsh(d[1]).max() <= o.eps_inf * nrm
                                   and np.linalg.eigvalsh(d[2]).max() <= o.eps_inf * nrm
                                   and float(d[3]) <= o.eps_inf * nrm and d[6].max() <= o.eps_inf * nrm)
                        if cone_ok and sup - bdy < -o.eps_inf * nrm:
                            status = STATUS_INFEASIBLE
                            break
            # ---- adaptive rho (residual balancing, OSQP style)
            if o.adaptive_rho and it % o.adapt_every == 0:
                ratio = np.sqrt((rp / max(n_p, 1e-12)) / max(rd / max(n_d, 1e-12), 1e-30))
                if ratio > o.adapt_thresh or ratio < 1.0 / o.adapt_thresh:
                    rho = float(np.clip(rho * ratio, 1e-6, 1e6))
    st.rho = rho
    X, Y, T, U = st.X, st.Y / c.a, st.T * c.a, st.U / c.sa
    obj = compute_SDP_relaxation_objective(X, Y, T, U, A, mask, gamma)
    # dual objective with lambda = -mu in K*:  -1/2 w'Pw - b'lambda + inf_{s in sets} lambda's + c0
    dual = (-0.5 * float(np.sum(Mk * st.X * st.X)) + c.c0
            + np.trace(st.m2[n:, n:]) + c.a * np.trace(st.m3) + c.ktr * st.m4 + float(c.beta @ st.mg)
            - float(np.sum(np.where(st.m5 < 0, st.m5 * c.lo, st.m5 * c.hi)))
            - float(np.sum(np.where(st.mv < 0, st.mv * c.lb, st.mv * c.ub))))
    return dict(status=status, feasible=status != STATUS_INFEASIBLE, objective=obj, dual_objective=dual,
                X=X, Y=Y, Theta=T, U=U, iters=it, res_p=res_p, res_d=res_d, state=st, rho=rho, consts=c,
                projections=None if trackers is None else (sum(t.n_lr for t in trackers), sum(t.n_full for t in trackers)),
                suspect=n_suspect, bound_max=bound_max, c0=c.c0)


def certificate(res, A, mask, gamma, k):
    """Algorithm-independent KKT check of a returned point (see module header).

    primal_*: most negative slack of each constraint of the ORIGINAL program at the returned point;
    dual_cone: largest positive eigenvalue / entry of mu (must lie in the polar cone);
    stationarity: ||P w + q - A' mu||_inf in the scaled program; gap: primal - dual objective.
    """
    n, m = A.shape
    X, Y, T, U = res["X"], res["Y"], res["Theta"], res["U"]
    st, c = res["state"], res["consts"]
    lo = c.lo / c.sa
    out = {}
    out["primal_psd1"] = min(0.0, np.linalg.eigvalsh(np.block([[Y, X], [X.T, T]])).min())
    out["primal_psd2"] = min(0.0, np.linalg.eigvalsh(np.block([[Y, U], [U.T, np.eye(k)]])).min())
    out["primal_psd3"] = min(0.0, np.linalg.eigvalsh(np.eye(n) - Y).min())
    out["primal_trace"] = min(0.0, k - np.trace(Y))
    out["primal_box"] = min(0.0, (U - lo).min(), (1.0 - U).min())
    out["primal_sym"] = max(np.abs(Y - Y.T).max(), np.abs(T - T.T).max())
    out["colnorm_minus_1"] = float(np.sqrt((U * U).sum(axis=0)).max() - 1.0)
    if c.L:
        v = c.x @ U
        agg = c.beta / c.a + np.sum(c.alpha / c.sa * v, axis=1) - np.einsum("li,ij,lj->l", c.x, Y, c.x)
        out["primal_cut_v"] = min(0.0, (v - c.lb / c.sa).min(), (c.ub / c.sa - v).min())
        out["primal_cut_agg"] = min(0.0, agg.min())
    out["dual_cone"] = max(0.0, np.linalg.eigvalsh(st.m1).max(), np.linalg.eigvalsh(st.m2).max(),
                           np.linalg.eigvalsh(st.m3).max(), st.m4, st.mg.max() if c.L else 0.0)
    gX, gY, gT, gU = c.At(st.m1, st.m2, st.m3, st.m4, st.m5, st.mv, st.mg)
    out["stationarity"] = max(np.abs(c.Mk * (st.X - A) - gX).max(), np.abs(gY).max(),
                              np.abs(c.cT * np.eye(m) - gT).max(), np.abs(gU).max())
    out["gap"] = res["objective"] - res["dual_objective"]
    return out
