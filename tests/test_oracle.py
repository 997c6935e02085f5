"""CPU tests: the oracle against analytic known answers, an independent second algorithm, KKT
certificates and brute-force restatements (the reference ships no golden vectors)."""
import itertools

import numpy as np
import pytest

from conftest import golden_instance, feasible_chain
from oracle import cuts as C
from oracle import eigsep as E
from oracle import kat
from oracle import mask as M
from oracle import objective as O
from oracle import relaxation as R
from oracle import shor as S
from oracle.datagen import config_instance, generate_matrix_completion_data

REL = 1e-6  # north_star: ~1e-6 relative on bounds at matched solver tolerance


def test_datagen_mask_has_exact_count_and_full_support():
    for seed in range(3):
        A, mask = generate_matrix_completion_data(1, 10, 10, 50, seed)
        assert mask.sum() == 50 and mask.any(axis=0).all() and mask.any(axis=1).all()
        assert A.shape == (10, 10)
    with pytest.raises(ValueError):
        generate_matrix_completion_data(1, 12, 10, 50, 0)   # n <= m, utils.jl:80-85


def test_objective_and_mse_against_loops():
    rng = np.random.default_rng(0)
    n, m, g = 7, 9, 80.0
    A = rng.standard_normal((n, m)); X = rng.standard_normal((n, m)); mask = rng.random((n, m)) < 0.5
    obj = sum(0.5 * (X[i, j] - A[i, j]) ** 2 for i in range(n) for j in range(m) if mask[i, j]) + (X ** 2).sum() / (2 * g)
    assert abs(O.evaluate_objective(X, A, mask, X[:, :2], g) - obj) <= 1e-13 * abs(obj)
    se = (X - A) ** 2
    assert abs(O.compute_MSE(X, A, mask, "in") - se[mask].mean()) < 1e-13
    assert abs(O.compute_MSE(X, A, mask, "out") - se[~mask].mean()) < 1e-13
    assert abs(O.compute_MSE(X, A, mask, "all") - se.mean()) < 1e-13
    # empty denominators return 0.0 (OMC.jl:2381-2382, 2390-2391)
    assert O.compute_MSE(X, A, np.ones_like(mask), "out") == 0.0
    assert O.compute_MSE(X, A, np.zeros_like(mask), "in") == 0.0
    with pytest.raises(ValueError):
        O.compute_MSE(X, A, mask, "bogus")
    T = rng.standard_normal((m, m))
    rel = O.compute_SDP_relaxation_objective(X, None, T, None, A, mask, g)
    assert abs(rel - (0.5 * ((A - X)[mask] ** 2).sum() + np.trace(T) / (2 * g))) < 1e-12


def test_bitmatrix_chunks_and_compaction_bit_exact():
    rng = np.random.default_rng(1)
    for n, m in [(1, 1), (3, 70), (10, 10), (13, 29), (64, 64)]:
        mask = rng.random((n, m)) < 0.4
        ch = M.bitmatrix_chunks(mask)
        assert ch.dtype == np.uint64 and ch.size == (n * m + 63) // 64
        for i, j in itertools.product(range(n), range(m)):
            b = i + n * j
            assert bool((int(ch[b >> 6]) >> (b & 63)) & 1) == bool(mask[i, j])
        assert np.array_equal(M.chunks_to_mask(ch, n, m), mask)
        rp, ci = M.mask_to_csr(mask); cp, ri = M.mask_to_csc(mask)
        assert rp[-1] == cp[-1] == mask.sum()
        for i in range(n):
            assert list(ci[rp[i]:rp[i + 1]]) == [j for j in range(m) if mask[i, j]]
        for j in range(m):
            assert list(ri[cp[j]:cp[j + 1]]) == [i for i in range(n) if mask[i, j]]


def test_cut_table_matches_reference_rows():
    """SURVEY appendix B / OMC.jl:1581-1676: every alpha v + beta is the secant of v^2 on [lb, ub],
    except linear3/right which reproduces the reference's expression |vhat| v (quirk Q1)."""
    for h in (-0.7, -0.2, 0.0, 0.35, 0.9):
        for t in C.LABELS:
            for d in C.LABELS[t]:
                lb, ub, al, be = C.cut_row(t, d, h)
                assert -1.0 <= lb <= ub <= 1.0
                if t == "linear2" and d == "middle":
                    assert al == 0.0 and be == h * h       # constant vhat^2, OMC.jl:1623
                    continue
                if t == "linear3" and d == "right":
                    assert (al, be) == (abs(h), 0.0)        # OMC.jl:1675
                    lb2, ub2, al2, be2 = C.cut_row(t, d, h, fix_linear3_right=True)
                    assert abs(al2 * lb2 + be2 - lb2 ** 2) < 1e-15 and abs(al2 * ub2 + be2 - ub2 ** 2) < 1e-15
                    continue
                assert abs(al * lb + be - lb * lb) < 1e-15 and abs(al * ub + be - ub * ub) < 1e-15


def test_child_enumeration_order_first_factor_fastest():
    """OMC.jl:2481-2491 + 2524: ind = 1 + sum_j code(dir_j) P^(j-1)."""
    for t, k in [("linear", 1), ("linear", 3), ("linear2", 2), ("linear3", 2)]:
        lab = C.LABELS[t]; P = len(lab)
        ch = C.child_directions(t, k)
        assert len(ch) == P ** k
        for ind, dirs in ch:
            assert ind == 1 + sum(lab.index(d) * P ** j for j, d in enumerate(dirs))
    assert C.child_directions("linear", 2) == [(1, ["left", "left"]), (2, ["right", "left"]),
                                               (3, ["left", "right"]), (4, ["right", "right"])]


def test_shor_indexes_against_brute_force_and_order():
    rng = np.random.default_rng(2)
    ind = rng.random((6, 7)) < 0.5
    for p in range(5):
        got = S.shor_constraint_indexes(ind, [p])
        assert len(got) == len(set(got)) and set(got) == S.shor_brute_force(ind, p)
        assert all(t[0] < t[1] and t[2] < t[3] for t in got)
    # the list follows the ORDER of num_entries_present_list (OMC.jl:2554)
    a = S.shor_constraint_indexes(ind, [4, 1]); b = S.shor_constraint_indexes(ind, [1, 4])
    n4 = len(S.shor_constraint_indexes(ind, [4]))
    assert a[:n4] == b[-n4:] and a[n4:] == b[:-n4]


def test_root_bound_closed_form_fully_observed():
    """KAT-root-full: closed-form water-filling value vs the ADMM oracle."""
    rng = np.random.default_rng(3)
    for (n, m, k) in [(6, 6, 1), (5, 8, 2), (10, 10, 1)]:
        A = rng.standard_normal((n, m))
        full = np.ones((n, m), bool)
        r = R.solve_relaxation(A, full, 80.0, k, opts=R.Options(eps_abs=1e-9, eps_rel=1e-9, max_iter=100000))
        want = kat.root_bound_full(A, 80.0, k)
        assert r["status"] == R.STATUS_OPTIMAL
        assert abs(r["objective"] - want) <= REL * want


def test_rank_k_optimum_pins_evaluate_objective():
    rng = np.random.default_rng(4)
    A = rng.standard_normal((7, 9)); full = np.ones_like(A, bool)
    for k in (1, 2, 3):
        X, obj = kat.rank_k_optimum_full(A, 80.0, k)
        assert abs(O.evaluate_objective(X, A, full, X[:, :k], 80.0) - obj) <= 1e-12 * obj
        assert kat.root_bound_full(A, 80.0, k) <= obj + 1e-12    # relaxation bound <= rank-k optimum


def test_root_bound_independent_projected_gradient():
    """Second algorithm (no cones, no ADMM) on PARTIALLY observed roots."""
    k, A, mask, g = config_instance("C1", 0)
    v, _ = kat.root_bound_projected_gradient(A, mask, g, k)
    r = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-9, eps_rel=1e-9))
    assert abs(r["objective"] - v) <= REL * v
    rng = np.random.default_rng(5)
    A2 = rng.standard_normal((6, 8)); m2 = rng.random((6, 8)) < 0.6; m2[0, :] = True; m2[:, 0] = True
    for kk in (1, 2):
        v, _ = kat.root_bound_projected_gradient(A2, m2, 80.0, kk)
        r = R.solve_relaxation(A2, m2, 80.0, kk, opts=R.Options(eps_abs=1e-9, eps_rel=1e-9))
        assert abs(r["objective"] - v) <= REL * v


def test_golden_cases_reproduce_and_are_kkt_certified(golden):
    """The committed golden vectors are regenerated by the oracle and carry an optimality certificate:
    primal feasible, dual feasible, zero gap -- independent of the algorithm that produced them."""
    for case in golden[:8]:
        A, mask, cuts = golden_instance(case)
        r = R.solve_relaxation(A, mask, case["gamma"], case["k"], case["cut_type"], cuts,
                               opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=100000))
        assert r["status"] == R.STATUS_OPTIMAL
        assert abs(r["objective"] - case["objective"]) <= REL * abs(case["objective"]), case["name"]
        cert = R.certificate(r, A, mask, case["gamma"], case["k"])
        for key, v in cert.items():
            if key.startswith("primal_") and key != "primal_sym":
                assert v >= -5e-6, (case["name"], key, v)
        assert cert["dual_cone"] <= 1e-9 and cert["stationarity"] <= 1e-5
        assert abs(cert["gap"]) <= 1e-4 * max(1.0, abs(case["objective"]))
        assert cert["colnorm_minus_1"] <= 1e-6          # the dropped ||U_j|| <= 1 rows hold (OMC.jl:1831-1835)
    for case in golden:
        c = case["certificate"]
        assert abs(c["gap"]) <= 1e-5 * max(1.0, abs(case["objective"])) and c["stationarity"] <= 1e-6


def test_children_bounds_dominate_parent_and_warm_start_agrees():
    k, A, mask, g = config_instance("C1", 0)
    o = R.Options(eps_abs=1e-8, eps_rel=1e-8)
    root = R.solve_relaxation(A, mask, g, k, opts=o)
    x, lam = E.breakpoint_vector(root["Y"], root["U"])
    assert lam[0] < -1e-6 and not E.master_feasible(root["Y"], root["U"])
    for _, dirs in C.child_directions("linear", k):
        cold = R.solve_relaxation(A, mask, g, k, "linear", [(x, root["U"], dirs)], opts=o)
        warm = R.solve_relaxation(A, mask, g, k, "linear", [(x, root["U"], dirs)], opts=o, state=root["state"])
        assert cold["objective"] >= root["objective"] - 1e-6
        assert abs(cold["objective"] - warm["objective"]) <= REL * cold["objective"]


def test_bound_invariant_to_sign_flip_with_directions_swapped():
    """x -> -x with left <-> right describes the same child."""
    k, A, mask, g = config_instance("C1", 1)
    o = R.Options(eps_abs=1e-8, eps_rel=1e-8)
    root = R.solve_relaxation(A, mask, g, k, opts=o)
    x, _ = E.breakpoint_vector(root["Y"], root["U"])
    rng = np.random.default_rng(0)
    Uh = 0.3 * rng.standard_normal(root["U"].shape)
    a = R.solve_relaxation(A, mask, g, k, "linear", [(x, Uh, ["left"])], opts=o)
    b = R.solve_relaxation(A, mask, g, k, "linear", [(-x, Uh, ["right"])], opts=o)
    assert abs(a["objective"] - b["objective"]) <= REL * a["objective"]


def test_separation_oracle_dense_matches_arpack():
    rng = np.random.default_rng(6)
    n, k = 30, 2
    U = rng.standard_normal((n, k)) / 6; B = rng.standard_normal((n, n)); Y = B @ B.T / n
    for nev in (1, 2):
        ld, vd = E.smallest_eigpairs(Y, U, nev, "dense")
        la, va = E.smallest_eigpairs(Y, U, nev, "arpack")
        assert np.allclose(ld, la, atol=1e-6)
        for q in range(nev):
            assert abs(abs(vd[:, q] @ va[:, q]) - 1) < 1e-4
            assert vd[np.argmax(np.abs(vd[:, q])), q] > 0          # sign rule
    x2, lam = E.breakpoint_vector(Y, U, "smallest_2_eigvec")
    w = np.abs(lam[:2]) / np.linalg.norm(lam[:2])
    assert lam[1] < -1e-10 and np.allclose(x2, w[0] * vd[:, 0] + w[1] * vd[:, 1])   # OMC.jl:2471-2473


def test_tracked_projection_follows_exact_projection():
    """oracle/lowrank.py (restatement of csrc/omc_lowrank.cuh): along a contracting sequence of symmetric matrices
    the tracked projection stays close to the exact one and needs the full eigensolver only to start."""
    from oracle import lowrank as LR
    rng = np.random.default_rng(0)
    N = 40
    Q, _ = np.linalg.qr(rng.standard_normal((N, N)))
    lam = np.concatenate([[3, 1, 0.2, 0.01], -np.abs(rng.standard_normal(N - 4))])
    V = (Q * lam) @ Q.T
    tp = LR.TrackedProjector(16)
    for it in range(80):
        D = rng.standard_normal((N, N)); D = (D + D.T) * 1e-3 * 0.9 ** it
        V = V + D
        P = tp.project(V)
        l, Qe = np.linalg.eigh(V)
        Pe = (Qe * np.maximum(l, 0)) @ Qe.T
        assert np.linalg.norm(P - Pe) <= 2.0 * np.linalg.norm(D) + 1e-12 * np.linalg.norm(V)   # error below the step itself
        assert np.linalg.eigvalsh(P).min() >= -1e-12                                           # always inside the cone
    assert tp.n_full == 1 and tp.n_lr == 79
    # negative side tracked when it is the smaller one; exact=True refreshes
    tq = LR.TrackedProjector(8)
    W = -V
    P0 = tq.project(W); assert tq.side == -1
    P1 = tq.project(W + 1e-6 * D)
    l, Qe = np.linalg.eigh(W + 1e-6 * D)
    assert np.linalg.norm(P1 - (Qe * np.maximum(l, 0)) @ Qe.T) <= 1e-8 * np.linalg.norm(W)


def test_admm_with_tracked_projection_reaches_the_same_bound():
    k, A, mask, g = config_instance("C1", 0)
    a = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-9, eps_rel=1e-9))
    b = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-9, eps_rel=1e-9, projection="tracked", pm=8))
    assert a["status"] == b["status"] == R.STATUS_OPTIMAL and abs(a["iters"] - b["iters"]) <= 26
    assert abs(a["objective"] - b["objective"]) <= 1e-8 * abs(a["objective"])
    lr, full = b["projections"]
    assert lr > full          # the 10 x 10 block is too small to track (2 p <= N): it stays on the full solver
    cert = R.certificate(b, A, mask, g, k)
    assert cert["dual_cone"] <= 1e-9 and cert["primal_psd1"] >= -1e-7 and abs(cert["gap"]) <= 1e-6


def test_tracked_projector_leaves_small_blocks_on_the_full_solver():
    """[Z R~] spans 2 p directions: a block with 2 (r + 2) > N is never tracked (found on the GPU with n = 2..5: the
    residual of a basis that spans the whole block is rounding noise and the Ritz step blew up)."""
    from oracle import lowrank as LR
    rng = np.random.default_rng(4)
    for N in (2, 3, 4, 5, 7):
        tp = LR.TrackedProjector(pm=16)
        V = rng.standard_normal((N, N)); V = V + V.T
        for _ in range(30):
            V = V + 1e-3 * (lambda E: E + E.T)(rng.standard_normal((N, N)))
            S = tp.project(V)
            assert np.abs(S - R.psd_project(0.5 * (V + V.T))).max() <= 1e-12
            assert tp.Z is None or 2 * tp.Z.shape[1] <= N
    k, A, mask = 1, *generate_matrix_completion_data(1, 4, 4, 12, 3)
    a = R.solve_relaxation(A, mask, 20.0, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8))
    b = R.solve_relaxation(A, mask, 20.0, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, projection="tracked"))
    assert a["status"] == b["status"] == R.STATUS_OPTIMAL and abs(a["objective"] - b["objective"]) <= 1e-8 * a["objective"]


def test_altmin_oracle_vstep_exact_ustep_matches_slsqp_and_stopping_rule():
    """oracle/altmin.py restates OMC.jl:1979-2279: the V-step is the exact minimiser in V (OMC.jl:2192-2209), the
    constrained U-step agrees with an independent SLSQP solve (OMC.jl:2212-2229 incl. the box, ball and pair-norm
    constraints), objectives[i] is the value after the i-th U-step and the stopping rule is OMC.jl:2232-2245."""
    from oracle import altmin as AM
    for cfg, eps_v in (("C1", 1e-9), ("C3", 1e-9)):
        k, A, mask, g = config_instance(cfg, 0)
        n, m = A.shape
        U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
        V = AM.v_step(U0, A, mask, g)
        f0 = AM.objective(U0, V, A, mask, g)
        rng = np.random.default_rng(3)
        for _ in range(5):                                    # exact minimiser of a strictly convex quadratic in V
            assert AM.objective(U0, V + 1e-4 * rng.standard_normal(V.shape), A, mask, g) > f0
        U1, f1, it, _ = AM.u_step(U0, V, A, mask, g)
        Us, fs = AM.ustep_slsqp(U0, V, A, mask, g)
        assert abs(f1 - fs) <= 1e-6 * abs(fs) and f1 <= f0 + 1e-12
        assert U1.max() <= 1 + 1e-8 and np.sqrt((U1 * U1).sum(axis=0)).max() <= 1 + 1e-7              # OMC.jl:2024, 2164-2171
        for j in range(k):
            assert U1[n - k + j:, j].min() >= -1e-8                                                        # OMC.jl:1989-1996
        r = AM.alternating_minimization(A, n, k, mask, g, True, "linear", U0)
        assert r["n_iters"] == len(r["objectives"]) <= r["max_iters"]
        ob = r["objectives"]
        if r["converged"] and len(ob) >= 2:
            rel = abs((ob[-1] - ob[-2]) / ob[-2])
            assert rel < 1e-5 or (len(ob) > 5 and all(ob[-1 - i] > ob[-6] for i in range(5)))


def test_package_generator_equals_oracle_generator():
    """bench.py feeds the engine from the package's own generator (the product never imports oracle/): same stream, same data."""
    from omc_b200 import synthetic as PS
    from oracle import datagen as OD
    for args in ((1, 10, 10, 50, 0), (2, 30, 30, 450, 1), (1, 50, 50, 1250, 0)):
        A1, m1 = PS.generate_matrix_completion_data(*args)
        A2, m2 = OD.generate_matrix_completion_data(*args)
        assert np.array_equal(A1, A2) and np.array_equal(m1, m2)


def test_oracle_certifies_an_infeasible_chain_in_exact_and_tracked_mode():
    """A 10-cut linear3 chain that the reference's `right` quirk (OMC.jl:1675, Q1) makes infeasible: the primal
    infeasibility certificate (d mu in the polar cone, A'd mu ~ 0, support < 0) stops it as INFEASIBLE (-> feasible =
    false, OMC.jl:1921-1935); the tracked mode switches the suspect node to exact projections and stops at the same
    check; with the valid secant (`fix_linear3_right`) the same chain is feasible."""
    n, m, k, ct, L = 6, 9, 2, "linear3", 10
    rng = np.random.default_rng(100 * n + 10 * k + L)
    A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
    cuts = feasible_chain(ct, n, k, L, rng)
    o = dict(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)
    ex = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(**o))
    tr = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(projection="tracked", **o))
    assert ex["status"] == tr["status"] == R.STATUS_INFEASIBLE and not ex["feasible"] and not tr["feasible"]
    assert tr["suspect"] == 1 and abs(tr["iters"] - ex["iters"]) <= 0.1 * ex["iters"]
    ok = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(fix_linear3_right=True, **o))
    assert ok["status"] == R.STATUS_OPTIMAL and ok["feasible"]


def test_oracle_infeasibility_by_bound_fires_early_and_never_on_feasible_nodes():
    """Options.infeasible_by_bound (kernel: -DOMC_INFEASIBLE_BY_BOUND, round-2 candidate): a feasible node has
    p* <= c0 = 1/2 ||P_Omega(A)||^2, so a certified lower bound above c0 proves infeasibility.  On the infeasible
    linear3 chain it fires at iteration 400 (d mu certificate: 5450); on feasible chains it never fires, the largest
    certified bound stays below the optimum and the iteration count is unchanged."""
    o = dict(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, infeasible_by_bound=True)
    n, m, k, ct, L = 6, 9, 2, "linear3", 10
    A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
    cuts = feasible_chain(ct, n, k, L, np.random.default_rng(100 * n + 10 * k + L))
    for proj in ("exact", "tracked"):
        r = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(projection=proj, **o))
        assert r["status"] == R.STATUS_INFEASIBLE and r["iters"] <= 500 and r["bound_max"] > r["c0"]
    for (n, m, k), ct, L in [((6, 9, 2), "linear2", 10), ((8, 8, 3), "linear", 10), ((4, 4, 1), "linear3", 3)]:
        A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
        cuts = feasible_chain(ct, n, k, L, np.random.default_rng(100 * n + 10 * k + L))
        r0 = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000))
        r1 = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(**o))
        assert r0["status"] == r1["status"] == R.STATUS_OPTIMAL and r0["iters"] == r1["iters"]
        assert r1["bound_max"] <= r0["objective"] * (1 + 1e-6) + 1e-9 <= r1["c0"] * (1 + 1e-6) + 1e-9
