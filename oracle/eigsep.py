"""Separation oracle and master-feasibility test.  TEST INFRASTRUCTURE ONLY.

Reference: ``eigs(Symmetric(U*U' - Y), nev=1|2, which=:SR, tol=1e-6)`` (ARPACK
through Arpack.jl 0.5.4, OMC.jl:2466-2477) and the feasibility test
``lambda_min(UU' - Y) >= -1e-6`` (OMC.jl:1272-1277).  ARPACK's eigenvector sign
depends on its random start vector; the engine fixes it deterministically
(largest-|.| component positive, first such index on ties), and so does this
restatement.  ``method="arpack"`` uses SciPy's ``eigsh(which="SA")`` -- the same
ARPACK routine family -- and ``method="dense"`` LAPACK ``eigh``.
"""
import numpy as np


def normalize_sign(x):
    i = int(np.argmax(np.abs(x)))
    return -x if x[i] < 0 else x


def smallest_eigpairs(Y, U, nev, method="dense"):
    M = U @ U.T - Y
    M = 0.5 * (M + M.T)
    if method == "arpack" and M.shape[0] > nev + 1:
        from scipy.sparse.linalg import eigsh
        lam, vec = eigsh(M, k=nev, which="SA", tol=1e-6)
        order = np.argsort(lam)
        lam, vec = lam[order], vec[:, order]
    else:
        lam, vec = np.linalg.eigh(M)
        lam, vec = lam[:nev], vec[:, :nev]
    vec = np.stack([normalize_sign(vec[:, i]) for i in range(nev)], axis=1)
    return lam, vec


def breakpoint_vector(Y, U, breakpoints="smallest_1_eigvec", method="dense"):
    """OMC.jl:2466-2477."""
    if breakpoints == "smallest_1_eigvec":
        lam, vec = smallest_eigpairs(Y, U, 1, method)
        return vec[:, 0], lam
    if breakpoints == "smallest_2_eigvec":
        lam, vec = smallest_eigpairs(Y, U, 2, method)
        if lam[1] < -1e-10:
            w = np.abs(lam[:2]) / np.sqrt(np.sum(lam[:2] ** 2))
            return w[0] * vec[:, 0] + w[1] * vec[:, 1], lam
        return vec[:, 0], lam
    raise ValueError("bad breakpoints")


def master_feasible(Y, U, projection_tolerance=1e-6, method="dense"):
    """OMC.jl:1272-1277 (disjunctive path)."""
    lam, _ = smallest_eigpairs(Y, U, 1, method)
    return bool(lam[0] >= -projection_tolerance)
