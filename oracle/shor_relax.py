"""CPU restatement of the node relaxation WITH Shor valid inequalities.  TEST INFRASTRUCTURE ONLY.
PARITY UNPINNED by the reference (no tests or golden vectors; Julia and Mosek are absent -- oracle/__init__.py): pinned by the
algorithm-independent properties listed at the end of this header.

The reference's program (OMC.jl:1491-1499 variables, 1503-1552 Shor variables, 1554-1561 main cones, 1564-1685 cut rows,
1755-1828 Shor rows, 1838-1846 objective), for any k:

    X = sum_t Xt[t]                                   (k > 1: OMC.jl:1492-1493; k = 1: Xt[1] = X)
    min 1/2 sum_I (A^2 - 2 A X + W) + tr(Theta)/(2 gamma)                                    OMC.jl:1838-1846
    s.t. [Y X; X' Theta] >= 0, [Y U; U' I] >= 0, I - Y >= 0, tr Y <= k, box on U, cut rows   (as oracle/relaxation.py)
         (1/2, W_ij, X_ij) in RSOC  for (i,j) in SOC_constraints_indexes                     OMC.jl:1757-1762, 1781-1786
         W_ij = sum_t Wt[t,ij] + 2 sum_{t1<t2} H[(t1,t2),ij]  on covered coordinates (k > 1)  OMC.jl:1787-1791
         Theta_jj = sum_i W_ij                                                                OMC.jl:1763-1767, 1792-1796
         per minor (i1,i2,j1,j2) and slice t:  the 5 x 5 moment matrix over (1, x11, x12, x21, x22) with diagonal Wt
             and the SAME variable V3 in the (x11,x22) and (x12,x21) positions >= 0          OMC.jl:1768-1779, 1797-1809
         per covered coordinate (k > 1): [1 Xt'; Xt diag(Wt) + offdiag(H)] >= 0              OMC.jl:1810-1826
         W >= 0, Wt >= 0                                                                      OMC.jl:1504, 1527-1530

Solved by the same conic ADMM (v-form, oracle/bigblock.py) with EXACT projections (eigh of every block), which small
instances can afford.  Elimination used (also by the CUDA engine): X and the covered W are not variables -- X = sum_t Xt,
W = sum Wt + 2 sum H -- so every cone row is a selection except the X part of the big block (sum over t), and the zero-cone
rows Theta_jj - sum_i W_ij = 0 (one per column j, disjoint supports): the w-update stays closed-form (per-coordinate and
per-column Sherman-Morrison).  Uncovered coordinates keep their own W (k = 1 semantics of the RSOC rows for any k).

Pins (tests/test_oracle_shor.py): no minors + all coordinates on RSOC rows == the plain relaxation (W = X^2 at the optimum);
Shor bound >= plain bound and <= the objective of any feasible rank-k point; zero duality gap.
"""
import itertools
import numpy as np
from .relaxation import Consts, Options, STATUS_OPTIMAL, STATUS_ITERATION_LIMIT, psd_project


def rsoc_project(v):
    """Projection of rows (a, b, x) onto the rotated second-order cone {2 a b >= x^2, a, b >= 0}."""
    a, b, x = v[..., 0], v[..., 1], v[..., 2]
    u = (a + b) / np.sqrt(2.0); w = (a - b) / np.sqrt(2.0)
    nr = np.sqrt(w * w + x * x)
    out_u = np.where(nr <= u, u, np.where(nr <= -u, 0.0, 0.5 * (u + nr)))
    scale = np.where(nr <= u, 1.0, np.where(nr <= -u, 0.0, 0.5 * (u + nr) / np.where(nr > 0, nr, 1.0)))
    w2, x2 = scale * w, scale * x
    return np.stack([(out_u + w2) / np.sqrt(2.0), (out_u - w2) / np.sqrt(2.0), x2], axis=-1)


class ShorStructure:
    """Index structure shared by all nodes: minors (0-based), covered / SOC coordinates, V1 / V2 variable numbering."""

    def __init__(self, n, m, k, minors, soc_coords):
        self.n, self.m, self.k = n, m, k
        self.minors = [tuple(int(v) for v in t) for t in minors]
        self.nm = len(self.minors)
        cov = np.zeros((n, m), bool)
        for (i1, i2, j1, j2) in self.minors:
            cov[i1, j1] = cov[i1, j2] = cov[i2, j1] = cov[i2, j2] = True
        self.covered = cov
        self.soc = np.zeros((n, m), bool)
        for (i, j) in soc_coords:
            self.soc[int(i), int(j)] = True
        assert not (self.soc & cov).any()
        self.cnt = np.zeros((n, m))                 # minors per coordinate
        self.v1_index, self.v2_index = {}, {}
        self.v1_of, self.v2_of = [], []             # per minor: (V1 id of row i1, of row i2), (V2 id of col j1, of col j2)
        for (i1, i2, j1, j2) in self.minors:
            for c in ((i1, j1), (i1, j2), (i2, j1), (i2, j2)):
                self.cnt[c] += 1
            a = self.v1_index.setdefault((i1, j1, j2), len(self.v1_index)); b = self.v1_index.setdefault((i2, j1, j2), len(self.v1_index))
            c_ = self.v2_index.setdefault((i1, i2, j1), len(self.v2_index)); d = self.v2_index.setdefault((i1, i2, j2), len(self.v2_index))
            self.v1_of.append((a, b)); self.v2_of.append((c_, d))
        self.nv1, self.nv2 = len(self.v1_index), len(self.v2_index)
        self.cnt1 = np.zeros(self.nv1); self.cnt2 = np.zeros(self.nv2)
        for (a, b), (c_, d) in zip(self.v1_of, self.v2_of):
            self.cnt1[a] += 1; self.cnt1[b] += 1; self.cnt2[c_] += 1; self.cnt2[d] += 1
        self.pairs = list(itertools.combinations(range(k), 2))
        self.mi = np.array(self.minors, dtype=int).reshape(-1, 4)


def _blocks5(S, Xt, Wd, V1, V2, V3):
    """(k, nm, 5, 5) moment matrices of the current point."""
    k = Xt.shape[0]
    B = np.zeros((k, S.nm, 5, 5))
    if S.nm == 0:
        return B
    i1, i2, j1, j2 = S.mi.T
    v1 = np.array(S.v1_of); v2 = np.array(S.v2_of)
    for t in range(k):
        x = [Xt[t, i1, j1], Xt[t, i1, j2], Xt[t, i2, j1], Xt[t, i2, j2]]
        w = [Wd[t, i1, j1], Wd[t, i1, j2], Wd[t, i2, j1], Wd[t, i2, j2]]
        B[t, :, 0, 0] = 1.0
        for s in range(4):
            B[t, :, 0, 1 + s] = x[s]; B[t, :, 1 + s, 0] = x[s]; B[t, :, 1 + s, 1 + s] = w[s]
        a, b = V1[t, v1[:, 0]], V1[t, v1[:, 1]]
        c_, d = V2[t, v2[:, 0]], V2[t, v2[:, 1]]
        e = V3[t]
        B[t, :, 1, 2] = B[t, :, 2, 1] = a      # (x11, x12): V1[i1, (j1, j2)]
        B[t, :, 1, 3] = B[t, :, 3, 1] = c_     # (x11, x21): V2[(i1, i2), j1]
        B[t, :, 1, 4] = B[t, :, 4, 1] = e      # (x11, x22): V3
        B[t, :, 2, 3] = B[t, :, 3, 2] = e      # (x12, x21): V3
        B[t, :, 2, 4] = B[t, :, 4, 2] = d      # (x12, x22): V2[(i1, i2), j2]
        B[t, :, 3, 4] = B[t, :, 4, 3] = b      # (x21, x22): V1[i2, (j1, j2)]
    return B


def _psd_batch(V):
    lam, Q = np.linalg.eigh(0.5 * (V + np.swapaxes(V, -1, -2)))
    return (Q * np.maximum(lam, 0.0)[..., None, :]) @ np.swapaxes(Q, -1, -2)


def solve_relaxation_shor(A, mask, gamma, k, minors, soc_coords, cut_type=None, cuts=(), opts=None):
    """One node with Shor rows.  ``minors``: 0-based (i1, i2, j1, j2); ``soc_coords``: 0-based (i, j) of the RSOC rows.
    Returns a dict like oracle/relaxation.solve_relaxation plus W and the dual objective."""
    o = opts or Options()
    A = np.asarray(A, dtype=float); mask = np.asarray(mask, dtype=bool)
    c = Consts(A, mask, gamma, k, cut_type, cuts, o)
    n, m, L = c.n, c.m, c.L
    S = ShorStructure(n, m, k, minors, soc_coords)
    Mk = c.Mk
    a_, sa = c.a, c.sa
    e9 = 1.0 if k > 1 else 0.0
    npair = len(S.pairs)
    # variables (scaled like oracle/relaxation.py: Y~ = aY, U~ = sqrt(a)U, Theta~ = Theta/a; X, W unscaled)
    Xt = np.zeros((k, n, m)); Wd = np.zeros((k, n, m)); H = np.zeros((npair, n, m))
    Y = np.zeros((n, n)); T = np.zeros((m, m)); U = np.zeros((n, k))
    V1 = np.zeros((k, S.nv1)); V2 = np.zeros((k, S.nv2)); V3 = np.zeros((k, S.nm))
    # v-form rows
    v1 = np.zeros((n + m, n + m)); v2 = c.E2.copy(); v3 = c.I3.copy(); v4 = c.ktr; v5 = np.zeros((n, k))
    vv = np.zeros((L, k)); vg = np.maximum(c.beta, 0.0) if L else np.zeros(0)
    if L:
        vv = np.clip(vv, c.lb, c.ub)
    vB = np.zeros((k, S.nm, 5, 5)); vB[:, :, 0, 0] = 1.0           # 5 x 5 blocks
    v9 = np.zeros((n, m, k + 1, k + 1)); v9[:, :, 0, 0] = 1.0      # (k+1) blocks on covered coordinates (k > 1)
    v6 = np.zeros(m)                                               # zero cone: Theta_jj - sum_i W_ij (scaled: a Theta~_jj)
    v7 = np.zeros((k, n, m))                                       # Wd >= 0
    vs = np.zeros((n, m, 3)); vs[:, :, 0] = 0.5                    # RSOC rows (1/2, W, X) on SOC coordinates
    rho = o.rho; sig, al = o.sigma, o.alpha
    G = c.gram(); r = G.shape[0]
    eyen, eyem = np.eye(n), np.eye(m)
    use9 = (k > 1)
    cov = S.covered.astype(float); socf = S.soc.astype(float)
    status = STATUS_ITERATION_LIMIT
    res_p = res_d = np.inf
    it = 0

    def Wfull(Wd_, H_):
        W = Wd_.sum(axis=0)
        if npair:
            W = W + 2.0 * H_.sum(axis=0)
        return W

    actW = np.stack([(S.covered | S.soc) if t == 0 else S.covered for t in range(k)]).astype(float)   # Wd[t] exists there
    actH = cov
    qX = -Mk * A                                                # objective -A X on Omega (linear in X = sum_t Xt)
    qW = 0.5 * Mk                                               # +1/2 W on Omega; W = sum Wd + 2 sum H

    def project_all():
        return dict(s1=psd_project(v1), s2=psd_project(v2), s3=psd_project(v3), s4=max(v4, 0.0), s5=np.clip(v5, c.lo, c.hi),
                    sv=(np.clip(vv, c.lb, c.ub) if L else vv), sg=np.maximum(vg, 0.0), sB=(_psd_batch(vB) if S.nm else vB),
                    s9=(_psd_batch(v9) if use9 else v9), s7=np.maximum(v7, 0.0), ss=rsoc_project(vs))

    def adjoint(t1, t2, t3, t4, t5, tv, tg, tB, t9, t7, ts, t6):
        """A' t for row values t (= -Sel' t for the selection rows), per variable block."""
        g = {}
        g["Y"] = -t1[:n, :n] - t2[:n, :n] + t3 + t4 * eyen + ((c.x.T * tg) @ c.x if L else 0.0)
        g["U"] = -(t2[:n, n:] + t2[n:, :n].T) - t5 - (c.x.T @ (tv + tg[:, None] * c.alpha) if L else 0.0)
        g["T"] = -t1[n:, n:].copy()
        g["T"][np.arange(m), np.arange(m)] += -a_ * t6                      # zero-cone rows: +a on Theta~_jj
        gXt = np.zeros((k, n, m)); gWd = np.zeros((k, n, m)); gH = np.zeros((npair, n, m))
        gV1 = np.zeros((k, S.nv1)); gV2 = np.zeros((k, S.nv2)); gV3 = np.zeros((k, S.nm))
        gXt += (-(t1[:n, n:] + t1[n:, :n].T))[None]                         # X part of the big block (row = sum_t Xt)
        gXt += (-ts[:, :, 2] * socf)[None]                                  # RSOC rows act on X = sum_t Xt
        if S.nm:
            i1, i2, j1, j2 = S.mi.T
            va = np.array(S.v1_of); vb = np.array(S.v2_of)
            for t in range(k):
                for s_, (ii, jj) in enumerate(((i1, j1), (i1, j2), (i2, j1), (i2, j2))):
                    np.add.at(gXt[t], (ii, jj), -2.0 * tB[t, :, 0, 1 + s_])
                    np.add.at(gWd[t], (ii, jj), -tB[t, :, 1 + s_, 1 + s_])
                np.add.at(gV1[t], va[:, 0], -2.0 * tB[t, :, 1, 2]); np.add.at(gV1[t], va[:, 1], -2.0 * tB[t, :, 3, 4])
                np.add.at(gV2[t], vb[:, 0], -2.0 * tB[t, :, 1, 3]); np.add.at(gV2[t], vb[:, 1], -2.0 * tB[t, :, 2, 4])
                gV3[t] = -2.0 * (tB[t, :, 1, 4] + tB[t, :, 2, 3])
        if use9:
            for t in range(k):
                gXt[t] += -2.0 * t9[:, :, 0, 1 + t] * cov
                gWd[t] += -t9[:, :, 1 + t, 1 + t] * cov
            for pi, (ta, tb_) in enumerate(S.pairs):
                gH[pi] = -2.0 * t9[:, :, 1 + ta, 1 + tb_] * cov
        gWd += -t7
        gWd[0] += -ts[:, :, 1] * socf
        gWd += t6[None, None, :]                                            # zero-cone rows: -1 on Wd
        gH += 2.0 * t6[None, None, :]                                       #                 -2 on H
        g.update(Xt=gXt, Wd=gWd * actW, H=gH * actH[None], V1=gV1, V2=gV2, V3=gV3)
        return g

    for it in range(1, o.max_iter + 1):
        P = project_all()
        s1, s2, s3, s4, s5, sv, sg, sB, s9, s7, ss = (P[q] for q in ("s1", "s2", "s3", "s4", "s5", "sv", "sg", "sB", "s9", "s7", "ss"))
        # ---- t / rho = v - 2 s (+ b)
        g = adjoint(v1 - 2 * s1, v2 - 2 * s2 + c.E2, v3 - 2 * s3 + c.I3, v4 - 2 * s4 + c.ktr, v5 - 2 * s5, vv - 2 * sv,
                    vg - 2 * sg + (c.beta if L else 0.0), vB - 2 * sB, v9 - 2 * s9, v7 - 2 * s7, vs - 2 * ss, v6)
        # ---- Y, U: as oracle/relaxation.py, dense-row Woodbury
        dYU = sig + 3.0 * rho
        Yt = (sig * Y + rho * g["Y"]) / dYU
        Ut = (sig * U + rho * g["U"]) / dYU
        cw = np.linalg.solve(np.eye(r) * (dYU / rho) + G, c.R(Yt, Ut))
        cY, cU = c.Rt(cw)
        Yt = Yt - cY; Ut = Ut - cU
        # ---- Xt: (sig + 2 rho (cnt + e9)) x_t + (2 rho + rho soc) sum_s x_s = rhs_t   (per-coordinate Sherman-Morrison)
        dX = sig + 2.0 * rho * (S.cnt + e9 * cov)
        cplX = 2.0 * rho + rho * socf
        rhsX = sig * Xt - qX[None] + rho * g["Xt"]
        Ssum = rhsX.sum(axis=0) / (dX + cplX * k)
        Xtt = (rhsX - cplX[None] * Ssum[None]) / dX[None]
        # ---- Wd, H, Theta~_jj: per-column Sherman-Morrison for the zero-cone rows  a Theta~_jj - sum_i W_ij = 0
        soc0 = np.zeros((k, n, m)); soc0[0] = socf
        dW = sig + rho * (S.cnt[None] + e9 * cov[None] + 1.0 + soc0)
        rW = (sig * Wd - qW[None] + rho * g["Wd"]) * actW
        dH = (sig + 2.0 * rho) * np.ones((npair, n, m)); rH = (sig * H - (2.0 * qW)[None] + rho * g["H"]) * actH[None]
        dTd = sig + rho
        rTd = sig * np.diag(T) - c.cT + rho * np.diag(g["T"])
        uDr = a_ * rTd / dTd - (rW / dW).sum(axis=(0, 1)) - 2.0 * ((rH / dH).sum(axis=(0, 1)) if npair else 0.0)
        uDu = a_ * a_ / dTd + (actW / dW).sum(axis=(0, 1)) + 4.0 * ((actH[None] / dH).sum(axis=(0, 1)) if npair else 0.0)
        coef = rho * uDr / (1.0 + rho * uDu)
        Wdt = (rW / dW + coef[None, None, :] / dW) * actW
        Ht = ((rH / dH + 2.0 * coef[None, None, :] / dH) * actH[None]) if npair else H
        Tdt = rTd / dTd - a_ * coef / dTd
        Tt = (sig * T - c.cT * eyem + rho * g["T"]) / (sig + rho)
        Tt[np.arange(m), np.arange(m)] = Tdt
        # ---- V1, V2, V3 (diagonal)
        V1t = (sig * V1 + rho * g["V1"]) / (sig + 2.0 * rho * S.cnt1[None]) if S.nv1 else V1
        V2t = (sig * V2 + rho * g["V2"]) / (sig + 2.0 * rho * S.cnt2[None]) if S.nv2 else V2
        V3t = (sig * V3 + rho * g["V3"]) / (sig + 4.0 * rho) if S.nm else V3
        # ---- z~ = rows of w~ ; v <- v + alpha (z~ - s) ; w <- alpha w~ + (1 - alpha) w
        def rows(Xt_, Wd_, H_, Y_, T_, U_, V1_, V2_, V3_):
            Xs_ = Xt_.sum(axis=0)
            z = dict(z1=np.block([[Y_, Xs_], [Xs_.T, T_]]), z2=np.block([[Y_, U_], [U_.T, np.eye(k)]]), z3=c.I3 - Y_, z4=c.ktr - np.trace(Y_), z5=U_)
            if L:
                zv = c.x @ U_
                z["zv"] = zv; z["zg"] = c.beta + np.sum(c.alpha * zv, axis=1) - np.einsum("li,ij,lj->l", c.x, Y_, c.x)
            z["zB"] = _blocks5(S, Xt_, Wd_, V1_, V2_, V3_) if S.nm else vB
            z9 = np.zeros_like(v9); z9[:, :, 0, 0] = 1.0
            if use9:
                for t in range(k):
                    z9[:, :, 0, 1 + t] = z9[:, :, 1 + t, 0] = Xt_[t]; z9[:, :, 1 + t, 1 + t] = Wd_[t]
                for pi, (ta, tb_) in enumerate(S.pairs):
                    z9[:, :, 1 + ta, 1 + tb_] = z9[:, :, 1 + tb_, 1 + ta] = H_[pi]
            z["z9"] = z9
            z["z6"] = a_ * np.diag(T_) - Wfull(Wd_, H_).sum(axis=0)
            z["z7"] = Wd_
            z["zs"] = np.stack([np.full((n, m), 0.5), Wd_[0], Xs_], axis=-1)
            return z
        z = rows(Xtt, Wdt, Ht, Yt, Tt, Ut, V1t, V2t, V3t)
        v1 = v1 + al * (z["z1"] - s1); v2 = v2 + al * (z["z2"] - s2); v3 = v3 + al * (z["z3"] - s3); v4 = v4 + al * (z["z4"] - s4)
        v5 = v5 + al * (z["z5"] - s5)
        if L:
            vv = vv + al * (z["zv"] - sv); vg = vg + al * (z["zg"] - sg)
        if S.nm:
            vB = vB + al * (z["zB"] - sB)
        if use9:
            v9 = v9 + al * ((z["z9"] - s9) * cov[:, :, None, None])
        v6 = v6 + al * z["z6"]
        v7 = v7 + al * ((z["z7"] - s7) * actW)
        vs = vs + al * ((z["zs"] - ss) * socf[:, :, None])
        Xt = al * Xtt + (1 - al) * Xt; Wd = al * Wdt + (1 - al) * Wd; H = al * Ht + (1 - al) * H
        Y = al * Yt + (1 - al) * Y; T = al * Tt + (1 - al) * T; U = al * Ut + (1 - al) * U
        V1 = al * V1t + (1 - al) * V1; V2 = al * V2t + (1 - al) * V2; V3 = al * V3t + (1 - al) * V3

        if it % o.check_every == 0 or it == o.max_iter:
            P = project_all()
            zc = rows(Xt, Wd, H, Y, T, U, V1, V2, V3)
            rp = max(np.abs(zc["z1"] - P["s1"]).max(), np.abs(zc["z2"] - P["s2"]).max(), np.abs(zc["z3"] - P["s3"]).max(),
                     abs(zc["z4"] - P["s4"]), np.abs(zc["z5"] - P["s5"]).max(), np.abs(zc["z6"]).max(), np.abs((zc["z7"] - P["s7"]) * actW).max(),
                     np.abs((zc["zs"] - P["ss"]) * socf[:, :, None]).max())
            if S.nm:
                rp = max(rp, np.abs(zc["zB"] - P["sB"]).max())
            if use9:
                rp = max(rp, np.abs((zc["z9"] - P["s9"]) * cov[:, :, None, None]).max())
            if L:
                rp = max(rp, np.abs(zc["zv"] - P["sv"]).max(), np.abs(zc["zg"] - P["sg"]).max())
            # dual residual: q - A'mu with mu = rho (v - s)  (P = 0: the Shor objective is linear)
            gm = adjoint(v1 - P["s1"], v2 - P["s2"], v3 - P["s3"], v4 - P["s4"], v5 - P["s5"], vv - P["sv"], vg - P["sg"],
                         vB - P["sB"], v9 - P["s9"], v7 - P["s7"], vs - P["ss"], v6)
            rd = max(np.abs(qX[None] - rho * gm["Xt"]).max(), np.abs((qW[None] - rho * gm["Wd"]) * actW).max(), np.abs(rho * gm["Y"]).max(),
                     np.abs(c.cT * eyem - rho * gm["T"]).max(), np.abs(rho * gm["U"]).max())
            if npair:
                rd = max(rd, np.abs(((2.0 * qW)[None] - rho * gm["H"]) * actH[None]).max())
            if S.nm:
                rd = max(rd, np.abs(rho * gm["V1"]).max(), np.abs(rho * gm["V2"]).max(), np.abs(rho * gm["V3"]).max())
            n_p = max(np.abs(P["s1"]).max(), c.a, c.ktr, 1.0)
            n_d = max(np.abs(qX).max(), c.cT, 1.0)
            res_p, res_d = rp, rd
            if o.verbose:
                print(f"it {it:6d} rp {rp:.3e} rd {rd:.3e} rho {rho:.3e}")
            if rp <= o.eps_abs + o.eps_rel * n_p and rd <= o.eps_abs + o.eps_rel * n_d:
                status = STATUS_OPTIMAL
                break
            if o.adaptive_rho and it % o.adapt_every == 0:
                ratio = np.sqrt((rp / n_p) / max(rd / n_d, 1e-30))
                if ratio > o.adapt_thresh or ratio < 1.0 / o.adapt_thresh:
                    rho_new = float(np.clip(rho * ratio, 1e-6, 1e6)); cf = rho / rho_new
                    resc = lambda v, s_: s_ + cf * (v - s_)
                    v1 = resc(v1, P["s1"]); v2 = resc(v2, P["s2"]); v3 = resc(v3, P["s3"]); v4 = resc(v4, P["s4"]); v5 = resc(v5, P["s5"])
                    if L:
                        vv = resc(vv, P["sv"]); vg = resc(vg, P["sg"])
                    if S.nm:
                        vB = resc(vB, P["sB"])
                    if use9:
                        v9 = resc(v9, P["s9"])
                    v6 = cf * v6; v7 = resc(v7, P["s7"]); vs = resc(vs, P["ss"])
                    rho = rho_new
    Xs = Xt.sum(axis=0); Wf = Wfull(Wd, H)
    Yo, To, Uo = Y / a_, T * a_, U / sa
    obj = 0.5 * float(np.sum(Mk * (A * A - 2.0 * A * Xs + Wf))) + float(np.trace(To)) / (2.0 * gamma)      # OMC.jl:1961-1968
    return dict(status=status, objective=obj, X=Xs, Xt=Xt, Y=Yo, Theta=To, U=Uo, W=Wf, Wd=Wd, H=H, iters=it, res_p=res_p, res_d=res_d,
                V1=V1, V2=V2, V3=V3, structure=S)
