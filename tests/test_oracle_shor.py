"""CPU tests of oracle/shor_relax.py: the node relaxation with Shor valid inequalities (OMC.jl:1503-1552, 1755-1846), restated
with exact projections.  The reference has no tests and Mosek is absent, so the restatement is pinned by properties that do
not depend on the algorithm:
  * without minors, every coordinate on an RSOC row, the Shor FORM of the program (W in place of X^2, Theta_jj = sum_i W_ij,
    linear objective) has the value of the plain relaxation whenever every column has an unobserved entry (W can then be
    inflated where it costs nothing); with a fully observed column it is strictly tighter;
  * with all 2 x 2 minors:  plain bound <= Shor bound <= objective of any rank-k point (the inequalities are valid)."""
import numpy as np

from oracle import relaxation as R
from oracle import shor as SI
from oracle import shor_relax as SR
from oracle.datagen import generate_matrix_completion_data


def _rank_k_local_optimum(A, mask, gamma, k, trials=12, sweeps=120):
    n, m = A.shape
    rng = np.random.default_rng(0)
    best = np.inf
    for _ in range(trials):
        Uf = rng.standard_normal((n, k)); Vf = rng.standard_normal((k, m))
        for _ in range(sweeps):
            for j in range(m):
                idx = mask[:, j]; Ui = Uf[idx]
                Vf[:, j] = np.linalg.solve(Ui.T @ Ui + Uf.T @ Uf / gamma + 1e-12 * np.eye(k), Ui.T @ A[idx, j])
            for i in range(n):
                idx = mask[i, :]; Vi = Vf[:, idx]
                Uf[i] = np.linalg.solve(Vi @ Vi.T + Vf @ Vf.T / gamma + 1e-12 * np.eye(k), Vi @ A[i, idx])
        X = Uf @ Vf
        best = min(best, 0.5 * ((X - A)[mask] ** 2).sum() + (X ** 2).sum() / (2 * gamma))
    return best


def test_rsoc_projection():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((200, 3)) * 2
    p = SR.rsoc_project(v)
    assert (2 * p[:, 0] * p[:, 1] >= p[:, 2] ** 2 - 1e-12).all() and (p[:, :2] >= -1e-12).all()     # in the cone
    assert np.abs(SR.rsoc_project(p) - p).max() <= 1e-12                                            # idempotent
    d = v - p                                                                                       # v - P(v) in the polar cone, orthogonal to P(v)
    assert np.abs((d * p).sum(axis=1)).max() <= 1e-10


def test_shor_form_without_minors_equals_plain_relaxation():
    for (k, n, m, nidx) in ((1, 6, 7, 22), (2, 6, 7, 22)):
        seed = next(s for s in range(1, 50) if not generate_matrix_completion_data(k, n, m, nidx, s)[1].all(axis=0).any())
        A, mask = generate_matrix_completion_data(k, n, m, nidx, seed)
        o = R.Options(eps_abs=1e-7, eps_rel=1e-7, max_iter=40000)
        plain = R.solve_relaxation(A, mask, 20.0, k, opts=o)
        r0 = SR.solve_relaxation_shor(A, mask, 20.0, k, [], [(i, j) for i in range(n) for j in range(m)], opts=o)
        assert r0["status"] == R.STATUS_OPTIMAL
        assert abs(r0["objective"] - plain["objective"]) <= 2e-5 * plain["objective"], (k, r0["objective"], plain["objective"])


def test_shor_bound_between_plain_bound_and_rank_k_optimum():
    for (k, n, m, nidx, seed) in ((1, 5, 6, 18, 3), (2, 5, 6, 20, 3)):
        A, mask = generate_matrix_completion_data(k, n, m, nidx, seed)
        o = R.Options(eps_abs=1e-7, eps_rel=1e-7, max_iter=40000)
        plain = R.solve_relaxation(A, mask, 20.0, k, opts=o)
        minors = [tuple(v - 1 for v in t) for t in SI.shor_constraint_indexes(mask, [1, 2, 3, 4])]
        cov = np.zeros((n, m), bool)
        for (i1, i2, j1, j2) in minors:
            cov[i1, j1] = cov[i1, j2] = cov[i2, j1] = cov[i2, j2] = True
        soc = [(i, j) for i in range(n) for j in range(m) if not cov[i, j]]
        r1 = SR.solve_relaxation_shor(A, mask, 20.0, k, minors, soc, opts=o)
        assert r1["status"] == R.STATUS_OPTIMAL
        ub = _rank_k_local_optimum(A, mask, 20.0, k)
        assert plain["objective"] * (1 - 2e-5) <= r1["objective"] <= ub * (1 + 1e-6), (k, plain["objective"], r1["objective"], ub)
        # the returned point satisfies the Shor rows: 5 x 5 moment blocks PSD, Theta_jj = sum_i W_ij, W >= 0
        S = r1["structure"]
        B = SR._blocks5(S, r1["Xt"], r1["Wd"], r1["V1"], r1["V2"], r1["V3"])
        assert np.linalg.eigvalsh(B).min() >= -1e-5
        assert np.abs(np.diag(r1["Theta"]) - r1["W"].sum(axis=0)).max() <= 1e-5 and r1["Wd"].min() >= -1e-6


def test_violated_minors_restatement():
    """Brute-force check of oracle/shor.py:violated_minors on a tiny case: a rank-1 slice scores 0 on every minor, a perturbed
    entry raises exactly the minors through it, existing minors are skipped, order is (score, tuple) descending."""
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(4), rng.standard_normal(5)
    X = np.outer(u, v)[None]
    cands = [(i1, i2, j1, j2) for i1 in range(4) for i2 in range(i1 + 1, 4) for j1 in range(5) for j2 in range(j1 + 1, 5)]
    top = SI.violated_minors(X, cands, [], 10)
    assert max(s for s, _ in top) <= 1e-14
    X2 = X.copy(); X2[0, 1, 2] += 0.5
    top = SI.violated_minors(X2, cands, [], len(cands))
    assert all((1 in (t[0], t[1]) and 2 in (t[2], t[3])) == (s > 1e-12) for s, t in top)
    assert [s for s, _ in top] == sorted((s for s, _ in top), reverse=True)
    skip = [t for _, t in top[:3]]
    top2 = SI.violated_minors(X2, cands, skip, 5)
    assert [t for _, t in top2] == [t for _, t in top[3:8]]
