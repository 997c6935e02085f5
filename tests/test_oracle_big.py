"""CPU tests of oracle/bigblock.py -- the NumPy restatement of the batched large-block engine (tracked, truncated-panel
projections, no eigendecomposition) -- against the exact-projection oracle (oracle/relaxation.py) and against the closed
form of the fully observed root (oracle/kat.py), at sizes the whole CPU suite can afford."""
import numpy as np
import pytest

from conftest import feasible_chain
from oracle import bigblock as Bg
from oracle import kat
from oracle import relaxation as R
from oracle.datagen import config_instance, generate_matrix_completion_data

REL = 1e-6


def test_tracked_engine_restatement_matches_exact_oracle_on_config_shapes():
    """Roots of the config 2 and config 3 shapes: same bound as the eigh-based oracle, about the same iteration count."""
    for cfg in ("C2", "C3"):
        k, A, mask, g = config_instance(cfg, 0)
        rb = Bg.solve_relaxation_big(A, mask, g, k, opts=Bg.BigOptions(eps_abs=1e-8, eps_rel=1e-8, max_iter=4000))
        re_ = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=4000))
        assert rb["status"] == R.STATUS_OPTIMAL and re_["status"] == R.STATUS_OPTIMAL
        assert abs(rb["objective"] - re_["objective"]) <= REL * abs(re_["objective"]), (cfg, rb["objective"], re_["objective"])
        assert rb["iters"] <= 1.3 * re_["iters"] + 50, (cfg, rb["iters"], re_["iters"])
        assert rb["lower_bound"] <= re_["objective"] * (1 + 1e-7) + 1e-9     # certified bound never above the optimum


def test_tracked_engine_restatement_with_cut_chains():
    """Cut chains of all three cut types (OMC.jl:1580-1683) on a k = 2 shape: bound equals the exact oracle's."""
    checked = 0
    for ct, L in (("linear", 3), ("linear2", 4), ("linear3", 3)):
        n, m, k = 12, 16, 2
        rng = np.random.default_rng(10 + L)
        A, mask = generate_matrix_completion_data(k, n, m, int(0.6 * n * m), 3)
        cuts = feasible_chain(ct, n, k, L, rng)
        re_ = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000))
        if re_["status"] != R.STATUS_OPTIMAL:
            continue
        rb = Bg.solve_relaxation_big(A, mask, 20.0, k, ct, cuts, opts=Bg.BigOptions(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000))
        assert rb["status"] == R.STATUS_OPTIMAL, (ct, L)
        assert abs(rb["objective"] - re_["objective"]) <= REL * abs(re_["objective"]), (ct, L, rb["objective"], re_["objective"])
        checked += 1
    assert checked >= 2


def test_tracked_engine_restatement_matches_closed_form_root_bound():
    """KAT-root-full: every entry observed, no cuts -> closed-form bound; sizes where the panel of 16
    is far narrower than the blocks (N = 140, 62, 60)."""
    for (n, m, k, seed) in ((60, 80, 2, 1), (40, 40, 3, 2)):
        A, _ = generate_matrix_completion_data(k, n, m, n * m, seed)
        mask = np.ones((n, m), bool)
        want = kat.root_bound_full(A, 80.0, k)
        rb = Bg.solve_relaxation_big(A, mask, 80.0, k, opts=Bg.BigOptions(eps_abs=1e-9, eps_rel=1e-9, max_iter=20000))
        assert rb["status"] == R.STATUS_OPTIMAL
        assert abs(rb["objective"] - want) <= REL * abs(want), (n, m, k, rb["objective"], want)


def test_probe_column_recovers_an_eigenvector_outside_the_trial_space():
    """Block-diagonal V whose second block is exactly orthogonal to the panel and to every residual: without the probe
    column the tracker can never see it (the failure met on [Y U; U' I] at a node without cuts)."""
    rng = np.random.default_rng(0)
    N1, N2 = 40, 3
    Q, _ = np.linalg.qr(rng.standard_normal((N1, N1)))
    V = np.zeros((N1 + N2, N1 + N2))
    V[:N1, :N1] = (Q * np.linspace(-5.0, 0.5, N1)) @ Q.T          # one small positive eigenvalue (0.5)
    V[N1:, N1:] = np.diag([3.0, 2.0, 1.0])                        # invisible block with the largest eigenvalues
    t = Bg.Tracker(N1 + N2, 16, +1, seed=5)
    t.Z[N1:, :] = 0.0                                             # panel entirely inside the first block
    t.Z = Bg.orthonormalize(t.Z)
    for q in range(12):
        t.step(V, tag=64 + 4 * q)
    lam = np.sort(np.linalg.eigvalsh(V))[::-1]
    assert np.abs(np.sort(t.th)[::-1][:4] - lam[:4]).max() <= 1e-8, (t.th, lam[:4])
