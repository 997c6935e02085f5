"""Objective / MSE restatements (plain arithmetic).  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def evaluate_objective(X, A, indices, U, gamma):
    """OMC.jl:2330-2359: 0.5*sum_I (X-A)^2 + ||X||_F^2/(2 gamma); U is only shape-checked."""
    if not (X.shape == A.shape == indices.shape and X.shape[0] == U.shape[0]):
        raise ValueError("Dimension mismatch.")
    d = (X - A)[indices]
    return 0.5 * float(np.dot(d, d)) + float(np.sum(X * X)) / (2.0 * gamma)


def compute_MSE(X, A, indices, kind="out"):
    """OMC.jl:2373-2409: masked / unmasked / overall MSE, 0.0 on an empty denominator set."""
    D2 = (X - A) ** 2
    total = indices.size
    nnz = int(indices.sum())
    if kind == "out":
        return 0.0 if total == nnz else float(D2[~indices].sum()) / (total - nnz)
    if kind == "in":
        return 0.0 if nnz == 0 else float(D2[indices].sum()) / nnz
    if kind == "all":
        return float(D2.sum()) / total
    raise ValueError('Input argument `kind` not recognized! Must be one of "out", "in", or "all".')


def compute_SDP_relaxation_objective(X, Y, Theta, U, A, indices, gamma,
                                     add_Shor_valid_inequalities=False, W=None):
    """OMC.jl:1945-1977: relaxation objective recomputed from the primal point."""
    tr = float(np.trace(Theta)) / (2.0 * gamma)
    if add_Shor_valid_inequalities:
        t = (A * A - 2.0 * A * X + W)[indices]
        return 0.5 * float(t.sum()) + tr
    d = (A - X)[indices]
    return 0.5 * float(np.dot(d, d)) + tr
