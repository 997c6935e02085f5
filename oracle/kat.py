"""Known-answer tests that do not depend on any conic solver.  TEST INFRASTRUCTURE ONLY.

They pin the oracle (and through it the CUDA engine) in place of the reference golden
vectors that do not exist (``/root/reference/test/runtests.jl`` is empty):

* ``root_bound_full`` -- with every entry observed and no cuts the relaxation
  OMC.jl:1554-1561,1848-1856 has the closed form  min 1/2 sum_i sigma_i^2 / (1 + gamma lambda_i)
  s.t. 0 <= lambda <= 1, sum lambda <= k  (eliminate Theta = X' Y^+ X, minimise over X column-wise
  to get 1/2 tr(A A' (I + gamma Y)^-1), then von Neumann's trace inequality).
* ``rank_k_optimum_full`` -- fully observed rank-constrained optimum X* = gamma/(1+gamma) A_k,
  objective 1/2 sum_{i>k} sigma_i^2 + 1/2 sum_{i<=k} sigma_i^2/(1+gamma)  (pins evaluate_objective
  and the value B&B must certify on fully observed inputs).
* ``root_bound_projected_gradient`` -- an independent second algorithm for the PARTIALLY observed
  root node: eliminating (X, Theta, U) leaves  min_{0 <= Y <= I, tr Y <= k} sum_j 1/2 a_j'(I + gamma
  Y_{I_j I_j})^-1 a_j  (U = 0 is feasible without cuts), a smooth convex problem solved here by
  projected gradient with exact spectral projections -- no ADMM, no cones.
"""
import numpy as np


def _capped_simplex(lam, k):
    """Euclidean projection of lam onto {0 <= x <= 1, sum x <= k}."""
    x = np.clip(lam, 0.0, 1.0)
    if x.sum() <= k:
        return x
    lo, hi = lam.min() - 1.0, lam.max()
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if np.clip(lam - mid, 0.0, 1.0).sum() > k:
            lo = mid
        else:
            hi = mid
    return np.clip(lam - hi, 0.0, 1.0)


def root_bound_full(A, gamma, k):
    sv = np.linalg.svd(A, compute_uv=False)
    f = lambda mu: np.clip((sv * np.sqrt(gamma / mu) - 1.0) / gamma, 0.0, 1.0)
    if f(1e-300).sum() <= k:
        lam = f(1e-300)
    else:
        lo, hi = 1e-30, 1e30
        for _ in range(600):
            mid = np.sqrt(lo * hi)
            if f(mid).sum() > k:
                lo = mid
            else:
                hi = mid
        lam = f(hi)
    return 0.5 * float(np.sum(sv ** 2 / (1.0 + gamma * lam)))


def rank_k_optimum_full(A, gamma, k):
    U, sv, Vt = np.linalg.svd(A, full_matrices=False)
    X = (gamma / (1.0 + gamma)) * (U[:, :k] * sv[:k]) @ Vt[:k]
    obj = 0.5 * float(np.sum(sv[k:] ** 2)) + 0.5 * float(np.sum(sv[:k] ** 2)) / (1.0 + gamma)
    return X, obj


def _reduced_objective(Y, A, mask, gamma, grad=False):
    n, m = A.shape
    val = 0.0
    G = np.zeros((n, n)) if grad else None
    for j in range(m):
        idx = np.flatnonzero(mask[:, j])
        if idx.size == 0:
            continue
        a = A[idx, j]
        M = np.eye(idx.size) + gamma * Y[np.ix_(idx, idx)]
        z = np.linalg.solve(M, a)
        val += 0.5 * float(a @ z)
        if grad:
            G[np.ix_(idx, idx)] -= 0.5 * gamma * np.outer(z, z)
    return (val, G) if grad else val


def root_bound_projected_gradient(A, mask, gamma, k, iters=4000, tol=1e-12):
    """Accelerated projected gradient (FISTA with function restart and backtracking) on the reduced problem."""
    n = A.shape[0]

    def proj(Y):
        lam, Q = np.linalg.eigh(0.5 * (Y + Y.T))
        return (Q * _capped_simplex(lam, k)) @ Q.T

    Y = proj(np.eye(n) * (k / n))
    Z, t, step = Y.copy(), 1.0, 1.0
    f_prev = _reduced_objective(Y, A, mask, gamma)
    for it in range(iters):
        fz, G = _reduced_objective(Z, A, mask, gamma, grad=True)
        while True:
            Yn = proj(Z - step * G)
            d = Yn - Z
            fn = _reduced_objective(Yn, A, mask, gamma)
            if fn <= fz + np.sum(G * d) + np.sum(d * d) / (2 * step) + 1e-15:
                break
            step *= 0.5
        tn = 0.5 * (1 + np.sqrt(1 + 4 * t * t))
        if fn > f_prev:            # restart
            Z, t = Y.copy(), 1.0
            f_prev = _reduced_objective(Y, A, mask, gamma)
            continue
        Z = Yn + ((t - 1) / tn) * (Yn - Y)
        if abs(f_prev - fn) <= tol * max(1.0, abs(fn)) and it > 50:
            Y, f_prev = Yn, fn
            break
        Y, t, f_prev = Yn, tn, fn
        step *= 1.5
    return f_prev, Y
