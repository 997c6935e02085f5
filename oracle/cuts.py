"""Disjunctive-cut coefficient table and child enumeration.  TEST INFRASTRUCTURE ONLY.

A cut is ``(x, Uhat, dirs)`` (OMC.jl:33-35).  Only ``vhat = Uhat' x`` is ever
read from ``Uhat`` (OMC.jl:1577, 2053).  For column ``j`` with ``h = vhat_j``
and ``a = |h|`` the relaxation adds ``lb <= v_j <= ub`` with ``v_j = x' U[:,j]``
and the aggregated row ``sum_j (alpha_j v_j + beta_j) >= x' Y x``
(OMC.jl:1580-1683).  ``linear3``/``right`` reproduces the reference's
expression ``|vhat| * v`` (OMC.jl:1675) unless
``fix_linear3_right`` is set.
"""
import itertools
import numpy as np

LABELS = {
    "linear": ["left", "right"],                                   # OMC.jl:2482
    "linear2": ["left", "middle", "right"],                        # OMC.jl:2486
    "linear3": ["left", "inner_left", "inner_right", "right"],     # OMC.jl:2490
}
TYPE_CODE = {"linear": 0, "linear2": 1, "linear3": 2}


def cut_row(cut_type, direction, h, fix_linear3_right=False):
    """Returns (lb, ub, alpha, beta) for one column of one cut."""
    a = abs(h)
    if cut_type == "linear":
        if direction == "left":           # OMC.jl:1582-1591
            return -1.0, h, h - 1.0, h
        if direction == "right":          # OMC.jl:1592-1601
            return h, 1.0, h + 1.0, -h
    elif cut_type == "linear2":
        if direction == "left":           # OMC.jl:1604-1613
            return -1.0, -a, -(1.0 + a), -a
        if direction == "middle":         # OMC.jl:1614-1623
            return -a, a, 0.0, h * h
        if direction == "right":          # OMC.jl:1624-1633
            return a, 1.0, 1.0 + a, -a
    elif cut_type == "linear3":
        if direction == "left":           # OMC.jl:1636-1645
            return -1.0, -a, -(1.0 + a), -a
        if direction == "inner_left":     # OMC.jl:1646-1655
            return -a, 0.0, -a, 0.0
        if direction == "inner_right":    # OMC.jl:1656-1665
            return 0.0, a, a, 0.0
        if direction == "right":          # OMC.jl:1666-1675
            if fix_linear3_right:
                return a, 1.0, 1.0 + a, -a
            return a, 1.0, a, 0.0
    raise ValueError(f"bad cut type/direction {cut_type}/{direction}")


def cut_rows(cut_type, cuts, fix_linear3_right=False):
    """Flatten a node's cut list to arrays: x[L,n], lb/ub/alpha[L,k], beta[L] (sum over j)."""
    L = len(cuts)
    if L == 0:
        return None
    n = cuts[0][0].shape[0]
    k = len(cuts[0][2])
    xs = np.zeros((L, n)); lb = np.zeros((L, k)); ub = np.zeros((L, k))
    al = np.zeros((L, k)); be = np.zeros(L)
    for l, (x, Uhat, dirs) in enumerate(cuts):
        vhat = Uhat.T @ x if Uhat.ndim == 2 else np.asarray(Uhat, dtype=float)
        xs[l] = x
        for j in range(k):
            lb[l, j], ub[l, j], al[l, j], b = cut_row(cut_type, dirs[j], float(vhat[j]), fix_linear3_right)
            be[l] += b
    return dict(x=xs, lb=lb, ub=ub, alpha=al, beta=be)


def child_directions(cut_type, k):
    """OMC.jl:2479-2493: ``enumerate(Iterators.product(repeat([labels], k)...))`` -- first factor fastest.

    Returns a list of (ind, dirs) with ind 1-based; child node_id = counter + ind (OMC.jl:2524).
    """
    labels = LABELS[cut_type]
    out = []
    for ind, rev in enumerate(itertools.product(labels, repeat=k), start=1):
        out.append((ind, list(rev[::-1])))  # itertools varies the LAST factor fastest; Julia the first
    return out


def direction_codes(cut_type, dirs):
    lab = LABELS[cut_type]
    return np.array([lab.index(d) for d in dirs], dtype=np.uint8)
