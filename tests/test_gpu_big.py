"""GPU parity tests of the batched large-block engine (engine = "batched", csrc/omc_big.cuh) through the C ABI, and the
algorithm-independent pins (closed-form root bound, projected-gradient solver) for BOTH relaxation engines.

Tolerance: 1e-6 relative on relaxation bounds (north_star); observed 1e-8 .. 1e-10."""
import numpy as np
import pytest

from conftest import feasible_chain as _feasible_chain

pytestmark = pytest.mark.gpu
REL = 1e-6


@pytest.fixture(scope="module")
def omc():
    import omc_b200
    omc_b200.init(0)
    return omc_b200


def _gc(omc, p, cuts):
    return [omc.Cut(p.add_cut(x, U), x, U, d) for x, U, d in cuts]


def _random_chain(n, k, L, ct, seed):
    from oracle.cuts import LABELS
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(L):
        x = rng.standard_normal(n); x /= np.linalg.norm(x)
        out.append((x, 0.3 * rng.standard_normal((n, k)), [LABELS[ct][rng.integers(len(LABELS[ct]) - 1)] for _ in range(k)]))
    return out


def test_batched_engine_matches_exact_oracle_on_config_shapes(omc):
    """Roots of the config 2, 3, 4 shapes (seeds 0, 1, 2 at config 2 / 3) and cut nodes of config 3 / 4: the batched engine's
    converged bound vs the eigh-based oracle, and its iteration count vs its own NumPy restatement."""
    from oracle import relaxation as R, bigblock as Bg
    from oracle.datagen import CONFIGS, generate_matrix_completion_data
    cases = [("C2", s, 0, 1e-8) for s in (0, 1, 2)] + [("C3", s, 0, 1e-8) for s in (0, 1, 2)] + [("C3", 0, 2, 1e-8), ("C4", 0, 0, 1e-9), ("C4", 0, 3, 1e-9)]
    for cfg, seed, L, eps in cases:
        c = CONFIGS[cfg]
        A, mask = generate_matrix_completion_data(c["k"], c["n"], c["m"], c["n_indices"], seed)
        cuts = _random_chain(c["n"], c["k"], L, c["cut_type"], 7) if L else []
        p = omc.Problem(c["k"], A, mask, 80.0, c["cut_type"])
        r = p.relax_batch([_gc(omc, p, cuts)], omc.default_opts(eps_abs=eps, eps_rel=eps, max_iter=8000), engine="batched")[0]
        ro = R.solve_relaxation(A, mask, 80.0, c["k"], c["cut_type"], cuts, opts=R.Options(eps_abs=eps, eps_rel=eps, max_iter=8000))
        assert ro["status"] == R.STATUS_OPTIMAL
        assert r["termination_status"] == "OPTIMAL", (cfg, seed, L, r["iters"])
        assert abs(r["objective"] - ro["objective"]) <= REL * abs(ro["objective"]), (cfg, seed, L, r["objective"], ro["objective"])
        assert r["lower_bound"] <= ro["objective"] * (1 + 1e-7) + 1e-9, (cfg, seed, L, r["lower_bound"], ro["objective"])
        assert r["iters"] <= 1.5 * ro["iters"] + 100, (cfg, seed, L, r["iters"], ro["iters"])
        # returned point is primal feasible for the reference's program (OMC.jl:1554-1561)
        n, k = c["n"], c["k"]
        Y, U = r["Y"], r["U"]
        assert np.abs(Y - Y.T).max() <= 1e-12
        assert np.linalg.eigvalsh(np.eye(n) - Y).min() >= -1e-6 and np.trace(Y) <= k + 1e-6
        assert np.linalg.eigvalsh(np.block([[Y, U], [U.T, np.eye(k)]])).min() >= -1e-6
        if cfg != "C4":
            rb = Bg.solve_relaxation_big(A, mask, 80.0, c["k"], c["cut_type"], cuts, opts=Bg.BigOptions(eps_abs=eps, eps_rel=eps, max_iter=8000))
            assert abs(r["iters"] - rb["iters"]) <= 0.15 * rb["iters"] + 50, (cfg, seed, L, r["iters"], rb["iters"])
        p.close()


def test_batched_engine_first_iteration_equals_the_restatement(omc):
    """The first lockstep iteration (node setup, cut table, Woodbury solve, the X / Theta and Y / U passes, the residual check)
    equals oracle/bigblock.py to rounding.  Later iterates are not compared one by one: while the minority side is wider than
    the panel the projection is a truncation, and which Ritz pair sits at the cut is sensitive to rounding (and the engine's
    Rayleigh-Ritz runs a few Jacobi sweeps where the restatement calls eigh); the converged bounds are compared instead."""
    from oracle import bigblock as Bg
    from oracle.datagen import config_instance
    for cfg, ct in (("C2", "linear"), ("C3", "linear2"), ("C4", "linear3")):
        k, A, mask, g = config_instance(cfg, 0)
        p = omc.Problem(k, A, mask, g, ct)
        cuts = _random_chain(A.shape[0], k, 2, ct, 3)
        gc = _gc(omc, p, cuts)
        r = p.relax_batch([gc], omc.default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=1, adapt_every=0), engine="batched")[0]
        ro = Bg.solve_relaxation_big(A, mask, g, k, ct, cuts, opts=Bg.BigOptions(eps_abs=1e-30, eps_rel=1e-30, max_iter=1, adaptive_rho=False))
        assert np.abs(r["X"] - ro["X"]).max() <= 1e-12 and np.abs(r["Y"] - ro["Y"]).max() <= 1e-12 and np.abs(r["U"] - ro["U"]).max() <= 1e-12
        # (the residuals hold the tracked factors after 6 steps: 3 Jacobi sweeps per Rayleigh-Ritz vs eigh in the restatement)
        assert abs(r["res_d"] - ro["res_d"]) <= 5e-2 * max(1.0, ro["res_d"])
        assert abs(r["res_p"] - ro["res_p"]) <= 5e-2 * max(1.0, ro["res_p"])
        p.close()


def test_batched_engine_batch_equals_singles_and_is_deterministic(omc):
    """A ragged batch (0..5 cuts per node, one node twice) relaxed in lockstep gives every node the result of a single-node
    run, bit for bit (no atomics on the data path; per-node control flow is independent of the batch)."""
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C3", 0)
    p = omc.Problem(k, A, mask, g, "linear2")
    sets = [[], _random_chain(30, 2, 1, "linear2", 1), _random_chain(30, 2, 2, "linear2", 2), _random_chain(30, 2, 5, "linear2", 3), []]
    gcs = [_gc(omc, p, cs) for cs in sets]
    o = omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=3000)
    rb = p.relax_batch(gcs, o, engine="batched")
    assert rb[0]["objective"] == rb[4]["objective"] and rb[0]["iters"] == rb[4]["iters"]
    for i, gc in enumerate(gcs):
        r1 = p.relax_batch([gc], o, engine="batched")[0]
        assert r1["iters"] == rb[i]["iters"] and r1["status_code"] == rb[i]["status_code"]
        assert r1["objective"] == rb[i]["objective"] and np.array_equal(r1["X"], rb[i]["X"])
    p.close()


def test_algorithm_independent_pins_closed_form_root_bound_both_engines(omc):
    """KAT-root-full (oracle/kat.py:root_bound_full): every entry observed, no cuts -> closed form
    min 1/2 sum sigma_i^2 / (1 + gamma lambda_i).  No ADMM, no cones, no eigensolver of ours: the strongest pin available
    without Mosek.  Config 1, 2, 4 sizes and a 300 x 400 block (N = 700) far beyond the persistent engine."""
    from oracle import kat
    from oracle.datagen import generate_matrix_completion_data
    for (k, n, m, engine) in ((1, 10, 10, "persistent"), (1, 50, 50, "persistent"), (1, 50, 50, "batched"), (2, 30, 30, "batched"),
                              (3, 100, 100, "batched"), (4, 300, 400, "batched")):
        A, _ = generate_matrix_completion_data(k, n, m, n * m, 1)
        mask = np.ones((n, m), bool)
        want = kat.root_bound_full(A, 80.0, k)
        p = omc.Problem(k, A, mask, 80.0, "linear")
        r = p.relax_batch([[]], omc.default_opts(eps_abs=1e-9, eps_rel=1e-9, max_iter=30000), engine=engine)[0]
        assert r["termination_status"] == "OPTIMAL", (k, n, m, engine, r["iters"])
        assert abs(r["objective"] - want) <= REL * abs(want), (k, n, m, engine, r["objective"], want)
        assert r["lower_bound"] <= want * (1 + 1e-7)
        p.close()


def test_algorithm_independent_pins_projected_gradient_both_engines(omc):
    """Partially observed roots against the projected-gradient solver of the reduced problem in Y (oracle/kat.py)."""
    from oracle import kat
    from oracle.datagen import generate_matrix_completion_data
    for (k, n, m, nidx, engine) in ((1, 10, 10, 50, "persistent"), (1, 10, 10, 50, "batched"), (2, 14, 18, 150, "persistent"), (2, 14, 18, 150, "batched")):
        A, mask = generate_matrix_completion_data(k, n, m, nidx, 2)
        want, _ = kat.root_bound_projected_gradient(A, mask, 80.0, k)
        p = omc.Problem(k, A, mask, 80.0, "linear")
        r = p.relax_batch([[]], omc.default_opts(eps_abs=1e-9, eps_rel=1e-9, max_iter=30000), engine=engine)[0]
        assert r["termination_status"] == "OPTIMAL"
        assert abs(r["objective"] - want) <= REL * abs(want), (k, n, m, engine, r["objective"], want)
        p.close()


def test_batched_engine_reports_infeasible_and_cutoff(omc):
    """The infeasible linear3 chain of test_gpu_parity (reference quirk Q1, OMC.jl:1675) comes back INFEASIBLE from the
    batched engine too (infeasibility by bound); a feasible node with an incumbent below its bound comes back CUTOFF."""
    from oracle import relaxation as R
    from oracle.datagen import generate_matrix_completion_data
    n, m, k, ct, L = 6, 9, 2, "linear3", 10
    rng = np.random.default_rng(100 * n + 10 * k + L)
    A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
    cuts = _feasible_chain(ct, n, k, L, rng)
    ro = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, infeasible_by_bound=True))
    assert ro["status"] == R.STATUS_INFEASIBLE
    p = omc.Problem(k, A, mask, 20.0, ct)
    gc = _gc(omc, p, cuts)
    r = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000), engine="batched")[0]
    assert r["termination_status"] == "INFEASIBLE" and not r["feasible"] and r["iters"] <= 3 * ro["iters"] + 100, (r["iters"], ro["iters"])
    fixed = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, fix_linear3_right=1), engine="batched")[0]
    rf = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, fix_linear3_right=True))
    assert fixed["termination_status"] == "OPTIMAL" and abs(fixed["objective"] - rf["objective"]) <= REL * rf["objective"]
    cut = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, fix_linear3_right=1, cutoff=0.5 * rf["objective"]), engine="batched")[0]
    assert cut["status_code"] == 4 and cut["objective"] > 0.5 * rf["objective"] and cut["iters"] < fixed["iters"]
    p.close()


def test_separation_oracle_beyond_the_in_sm_eigensolver(omc):
    """n > 104: restarted Lanczos (csrc/omc_big.cuh: k_lanczos) against dense eigh: smallest one / two eigenpairs of
    U U' - Y (OMC.jl:2466-2477), the mixed breakpoint vector (OMC.jl:2471-2476) and the feasibility bit (OMC.jl:1274-1276)."""
    rng = np.random.default_rng(11)
    for n, k, B in ((150, 3, 3), (400, 5, 2)):
        Ys, Us = [], []
        for b in range(B):
            Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
            lam = np.concatenate([rng.uniform(0.5, 1.0, k + 2), rng.uniform(0.0, 0.05, n - k - 2)])
            Ys.append((Q * lam) @ Q.T)
            Us.append(0.3 * rng.standard_normal((n, k)) / np.sqrt(n))
        Ys[-1] = Us[-1] @ Us[-1].T + 1e-9 * np.eye(n)            # master-feasible node: lambda_min(UU' - Y) = -1e-9
        Y = np.stack(Ys); U = np.stack(Us)
        for nev in (1, 2):
            lam_g, vec_g, bp_g, feas_g = omc.smallest_eigvecs_batch(Y, U, nev)
            for b in range(B):
                M = U[b] @ U[b].T - Y[b]
                w, V = np.linalg.eigh(0.5 * (M + M.T))
                assert abs(lam_g[b, 0] - w[0]) <= 1e-8 * max(1.0, abs(w[0])), (n, nev, b, lam_g[b], w[:2])
                assert feas_g[b] == (w[0] >= -1e-6)
                if b < B - 1:                                    # (the feasible node's spectrum is a flat cluster at ~0)
                    assert abs(abs(vec_g[b, :, 0] @ V[:, 0]) - 1.0) <= 1e-6
                    assert vec_g[b, np.abs(vec_g[b, :, 0]).argmax(), 0] > 0
                    if nev == 2:
                        assert abs(lam_g[b, 1] - w[1]) <= 1e-7 * max(1.0, abs(w[1]))
                        wt = np.abs(w[:2]) / np.linalg.norm(w[:2])
                        v0 = V[:, 0] * np.sign(V[np.abs(V[:, 0]).argmax(), 0]); v1 = V[:, 1] * np.sign(V[np.abs(V[:, 1]).argmax(), 1])
                        assert np.abs(bp_g[b] - (wt[0] * v0 + wt[1] * v1)).max() <= 1e-5
                    else:
                        assert np.abs(bp_g[b] - vec_g[b, :, 0]).max() == 0.0
