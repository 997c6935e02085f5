"""CPU restatement of alternating_minimization (OMC.jl:1979-2279).  TEST INFRASTRUCTURE ONLY.

f(U, V) = 1/2 sum_I ((UV)_ij - A_ij)^2 + 1/(2 gamma) ||UV||_F^2          OMC.jl:2196-2206, 2216-2226
  V-step: unconstrained, closed form per column (OMC.jl:2192-2209)
  U-step: min_U f(U, V) s.t. (OMC.jl:2021-2171)
            lo <= U <= 1 (symmetry-breaking zeros, OMC.jl:2024-2025, defaults 1989-1996)
            ||U_a + U_b|| <= sqrt 2, ||U_a - U_b|| <= sqrt 2  for a < b        OMC.jl:2030-2045
            lb <= x' U_j <= ub per cut and column (NOT the aggregated row)   OMC.jl:2049-2091
            ||U_j|| <= 1                                                    OMC.jl:2164-2171
  stop: |f_new - f_prev| / |f_prev| < eps (f_prev starts at 1e10), or more than 5 objectives and each of the
        last five exceeds objectives[end-5]; max_iters = 100              OMC.jl:2012, 2232-2245

The reference solves both half-steps with Mosek; here the V-step is the exact k x k normal equations and
the U-step is solved by ADMM in OSQP form (z = A u in C, C a product of boxes, balls and intervals), the
same algorithm the CUDA kernel runs.  ``ustep_slsqp`` is an independent SciPy solve used by the tests.
"""
import numpy as np
from .cuts import cut_rows
from .relaxation import default_U_lower


def objective(U, V, A, mask, gamma):
    X = U @ V
    d = (X - A)[mask]
    return 0.5 * float(d @ d) + float(np.sum(X * X)) / (2.0 * gamma)


def v_step(U, A, mask, gamma):
    n, k = U.shape
    m = A.shape[1]
    V = np.zeros((k, m))
    UtU = U.T @ U / gamma
    for j in range(m):
        idx = np.flatnonzero(mask[:, j])
        Uj = U[idx]
        H = Uj.T @ Uj + UtU
        g = Uj.T @ A[idx, j]
        try:
            V[:, j] = np.linalg.solve(H, g)
        except np.linalg.LinAlgError:
            V[:, j] = np.linalg.lstsq(H, g, rcond=None)[0]
    return V


def _ball(v, radius):
    nr = np.linalg.norm(v)
    return v if nr <= radius else v * (radius / nr)


class UStepState:
    def __init__(self, n, k, L):
        self.zb = np.zeros((n, k)); self.yb = np.zeros((n, k))            # box
        self.zc = np.zeros((n, k)); self.yc = np.zeros((n, k))            # column balls
        npair = k * (k - 1) // 2
        self.zp = np.zeros((npair, n)); self.yp = np.zeros((npair, n))    # U_a + U_b
        self.zm = np.zeros((npair, n)); self.ym = np.zeros((npair, n))    # U_a - U_b
        self.zv = np.zeros((L, k)); self.yv = np.zeros((L, k))            # cut rows
        self.rho = 1.0


def u_step(U0, V, A, mask, gamma, rows=None, st=None, eps=1e-9, max_iter=20000, sigma=1e-6, alpha=1.6):
    """ADMM for the constrained U-step.  Returns (U, objective value f(U, V), iterations, state)."""
    n, k = U0.shape
    L = 0 if rows is None else rows["x"].shape[0]
    lo, hi = default_U_lower(n, k), np.ones((n, k))
    pairs = [(a, b) for a in range(k - 1) for b in range(a + 1, k)]
    r2 = np.sqrt(2.0)
    VVt = V @ V.T / gamma
    H = np.zeros((n, k, k)); g = np.zeros((n, k))
    for i in range(n):
        idx = np.flatnonzero(mask[i])
        Vi = V[:, idx]
        H[i] = Vi @ Vi.T + VVt
        g[i] = Vi @ A[i, idx]
    st = st or UStepState(n, k, L)
    if st.zv.shape[0] != L:
        st.zv = np.zeros((L, k)); st.yv = np.zeros((L, k))
    U = U0.copy()
    x = rows["x"] if L else np.zeros((0, n))
    rho = st.rho
    K = None
    it = 0
    for it in range(1, max_iter + 1):
        if K is None:
            c = sigma + 2.0 * k * rho
            K = np.linalg.inv(H + c * np.eye(k)[None])                       # n x k x k
            if L:
                # M[(l,j),(l',j')] = sum_i x_l[i] x_l'[i] K_i[j,j'] + delta / rho
                M = np.einsum("li,mi,ijq->ljmq", x, x, K).reshape(L * k, L * k) + np.eye(L * k) / rho
                Minv = np.linalg.inv(M)
        # rhs = sigma u - q + A'(rho z - y)
        rhs = sigma * U + g + (rho * st.zb - st.yb) + (rho * st.zc - st.yc)
        for p, (a, b) in enumerate(pairs):
            tp = rho * st.zp[p] - st.yp[p]; tm = rho * st.zm[p] - st.ym[p]
            rhs[:, a] += tp + tm
            rhs[:, b] += tp - tm
        if L:
            rhs += x.T @ (rho * st.zv - st.yv)
        Ut = np.einsum("ijq,iq->ij", K, rhs)
        if L:
            cw = Minv @ (x @ Ut).reshape(-1)
            Ut = Ut - np.einsum("ijq,iq->ij", K, x.T @ cw.reshape(L, k))
        # z~ = A u~ ; relax ; project ; dual update
        Un = alpha * Ut + (1 - alpha) * U

        def upd(zt, z, y, proj):
            v = alpha * zt + (1 - alpha) * z + y / rho
            zn = proj(v)
            return zn, y + rho * (alpha * zt + (1 - alpha) * z - zn)

        st.zb, st.yb = upd(Ut, st.zb, st.yb, lambda v: np.clip(v, lo, hi))
        zc = np.zeros_like(st.zc); yc = np.zeros_like(st.yc)
        for j in range(k):
            zc[:, j], yc[:, j] = upd(Ut[:, j], st.zc[:, j], st.yc[:, j], lambda v: _ball(v, 1.0))
        st.zc, st.yc = zc, yc
        for p, (a, b) in enumerate(pairs):
            st.zp[p], st.yp[p] = upd(Ut[:, a] + Ut[:, b], st.zp[p], st.yp[p], lambda v: _ball(v, r2))
            st.zm[p], st.ym[p] = upd(Ut[:, a] - Ut[:, b], st.zm[p], st.ym[p], lambda v: _ball(v, r2))
        if L:
            st.zv, st.yv = upd(x @ Ut, st.zv, st.yv, lambda v: np.clip(v, rows["lb"], rows["ub"]))
        U = Un
        if it % 10 == 0 or it == max_iter:
            rp = max(np.abs(U - st.zb).max(), np.abs(U - st.zc).max(),
                     max([np.abs(U[:, a] + U[:, b] - st.zp[p]).max() for p, (a, b) in enumerate(pairs)] + [0.0]),
                     max([np.abs(U[:, a] - U[:, b] - st.zm[p]).max() for p, (a, b) in enumerate(pairs)] + [0.0]),
                     np.abs(x @ U - st.zv).max() if L else 0.0)
            gr = np.einsum("ijq,iq->ij", H, U) - g + st.yb + st.yc
            for p, (a, b) in enumerate(pairs):
                gr[:, a] += st.yp[p] + st.ym[p]
                gr[:, b] += st.yp[p] - st.ym[p]
            if L:
                gr += x.T @ st.yv
            rd = np.abs(gr).max()
            npr = max(1.0, np.abs(U).max()); ndr = max(1.0, np.abs(g).max(), np.abs(np.einsum("ijq,iq->ij", H, U)).max())
            if rp <= eps * npr and rd <= eps * ndr:
                break
            if it % 50 == 0:
                ratio = np.sqrt((rp / npr) / max(rd / ndr, 1e-30))
                if ratio > 5.0 or ratio < 0.2:
                    rho = float(np.clip(rho * ratio, 1e-6, 1e6)); K = None
    st.rho = rho
    return U, objective(U, V, A, mask, gamma), it, st


def ustep_slsqp(U0, V, A, mask, gamma, rows=None):
    """Independent solve of the same U-step with SciPy SLSQP (small instances only)."""
    from scipy.optimize import minimize
    n, k = U0.shape
    lo = default_U_lower(n, k)

    def f(u):
        U = u.reshape(n, k)
        return objective(U, V, A, mask, gamma)

    def grad(u):
        U = u.reshape(n, k)
        X = U @ V
        G = (mask * (X - A)) @ V.T + (X @ V.T) / gamma
        return G.reshape(-1)

    cons = []
    for j in range(k):
        cons.append({"type": "ineq", "fun": lambda u, j=j: 1.0 - np.sum(u.reshape(n, k)[:, j] ** 2)})
    for a in range(k - 1):
        for b in range(a + 1, k):
            cons.append({"type": "ineq", "fun": lambda u, a=a, b=b: 2.0 - np.sum((u.reshape(n, k)[:, a] + u.reshape(n, k)[:, b]) ** 2)})
            cons.append({"type": "ineq", "fun": lambda u, a=a, b=b: 2.0 - np.sum((u.reshape(n, k)[:, a] - u.reshape(n, k)[:, b]) ** 2)})
    if rows is not None:
        for l in range(rows["x"].shape[0]):
            for j in range(k):
                cons.append({"type": "ineq", "fun": lambda u, l=l, j=j: rows["x"][l] @ u.reshape(n, k)[:, j] - rows["lb"][l, j]})
                cons.append({"type": "ineq", "fun": lambda u, l=l, j=j: rows["ub"][l, j] - rows["x"][l] @ u.reshape(n, k)[:, j]})
    res = minimize(f, np.clip(U0, lo, 1.0).reshape(-1), jac=grad, bounds=list(zip(lo.reshape(-1), np.ones(n * k))),
                   constraints=cons, method="SLSQP", options=dict(maxiter=500, ftol=1e-14))
    return res.x.reshape(n, k), float(res.fun)


def alternating_minimization(A, n, k, indices, gamma, use_disjunctive_cuts=True, disjunctive_cuts_type=None, U_initial=None,
                             disjunctive_cuts=(), eps=1e-5, max_iters=100, inner_eps=1e-9, fix_linear3_right=False):
    """OMC.jl:1979-2279 (disjunctive path).  Returns the reference's result dict (OMC.jl:2249-2278)."""
    A = np.asarray(A, float); mask = np.asarray(indices, bool)
    rows = cut_rows(disjunctive_cuts_type, list(disjunctive_cuts), fix_linear3_right) if len(disjunctive_cuts) else None
    U = np.array(U_initial, float)
    V = np.zeros((k, A.shape[1]))
    objective_current = 1e10
    objectives, converged, counter, st = [], False, 0, None
    inner = 0
    while counter < max_iters:
        counter += 1
        V = v_step(U, A, mask, gamma)
        U, obj, it, st = u_step(U, V, A, mask, gamma, rows, st, eps=inner_eps)
        inner += it
        objectives.append(obj)
        if abs((obj - objective_current) / objective_current) < eps:
            converged = True
        elif len(objectives) > 5 and all(objectives[-1 - i] > objectives[-6] for i in range(5)):
            converged = True
        if converged:
            break
        objective_current = obj
    return {"converged": converged, "U": U, "V": V, "n_iters": counter, "max_iters": max_iters,
            "objectives": objectives, "inner_iters": inner}
