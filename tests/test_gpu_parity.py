"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle, the committed golden
vectors and size-independent properties.  Bit-exact for integer work; 1e-6 relative on relaxation bounds
(north_star), 1e-13 relative on objective/MSE reductions (summation order differs)."""
import numpy as np
import pytest

from conftest import golden_instance, feasible_chain as _feasible_chain

pytestmark = pytest.mark.gpu
REL_BOUND = 1e-6


@pytest.fixture(scope="module")
def omc():
    import omc_b200
    omc_b200.init(0)
    return omc_b200


def _cuts_for(omc, p, cuts):
    return [omc.Cut(p.add_cut(x, U), x, U, d) for x, U, d in cuts]


def test_eigensolver_psd_projection(omc):
    rng = np.random.default_rng(0)
    for N, B in [(1, 2), (2, 3), (11, 4), (20, 8), (51, 8), (64, 4), (100, 16), (104, 2)]:
        V = rng.standard_normal((B, N, N)); V = V + np.transpose(V, (0, 2, 1))
        V[0] = 0.0                                          # zero matrix
        if B > 1:
            V[1] = np.diag(rng.standard_normal(N))          # already diagonal
        P, lam, sw, _ = omc.psd_project_batch(V)
        for b in range(B):
            l, Q = np.linalg.eigh(V[b])
            assert np.abs(P[b] - (Q * np.maximum(l, 0)) @ Q.T).max() <= 1e-12 * max(1.0, np.abs(V[b]).max())
            assert np.abs(np.sort(lam[b]) - l).max() <= 1e-12 * max(1.0, np.abs(l).max())
    # idempotence and clustered spectra
    B, N = 4, 60
    Q, _ = np.linalg.qr(rng.standard_normal((N, N)))
    lam0 = np.concatenate([np.full(20, 3.0), np.full(20, -2.0), np.zeros(10), rng.standard_normal(10) * 1e-9])
    V = np.stack([(Q * lam0) @ Q.T] * B)
    P, _, _, _ = omc.psd_project_batch(V)
    P2, _, _, _ = omc.psd_project_batch(P)
    assert np.abs(P2 - P).max() <= 1e-12 and np.abs(P[0] - (Q * np.maximum(lam0, 0)) @ Q.T).max() <= 1e-12


def test_mask_compaction_bit_exact(omc):
    from oracle import mask as M
    rng = np.random.default_rng(1)
    for n, m, dens in [(1, 1, 1.0), (3, 70, 0.3), (10, 10, 0.5), (13, 29, 0.1), (50, 50, 0.5), (64, 64, 0.9), (40, 97, 0.0)]:
        mask = rng.random((n, m)) < dens
        p = omc.Problem(1, rng.standard_normal((n, m)), mask, 80.0)
        rp, ci, cp, ri = p.csr()
        orp, oci = M.mask_to_csr(mask); ocp, ori = M.mask_to_csc(mask)
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(cp, ocp) and np.array_equal(ri, ori)
        p.close()


def test_objective_mse_fused_reduction(omc):
    from oracle import objective as O
    rng = np.random.default_rng(2)
    for n, m, dens in [(1, 1, 1.0), (10, 10, 0.5), (50, 50, 0.5), (33, 77, 0.2), (20, 20, 1.0), (20, 20, 0.0), (1000, 1000, 0.2)]:
        A = rng.standard_normal((n, m)); X = rng.standard_normal((n, m)); mask = rng.random((n, m)) < dens
        p = omc.Problem(1, A, mask, 80.0)
        got = p.objective_mse(X)
        want = (O.evaluate_objective(X, A, mask, X[:, :1], 80.0), O.compute_MSE(X, A, mask, "in"),
                O.compute_MSE(X, A, mask, "out"), O.compute_MSE(X, A, mask, "all"))
        for g, w in zip(got, want):
            assert abs(g - w) <= 1e-13 * max(1.0, abs(w))
        assert omc.evaluate_objective(p, X) == got[0] and omc.compute_MSE(p, X, "in") == got[1]
        with pytest.raises(ValueError):
            omc.compute_MSE(p, X, "bogus")
        p.close()


def test_relaxation_matches_golden_vectors(omc, golden):
    for case in golden:
        A, mask, cuts = golden_instance(case)
        p = omc.Problem(case["k"], A, mask, case["gamma"], case["cut_type"])
        r = omc.matrix_completion_SDP_relaxation(p, _cuts_for(omc, p, cuts), omc.default_opts(eps_abs=1e-9, eps_rel=1e-9, max_iter=200000))
        assert r["termination_status"] == "OPTIMAL" and r["feasible"], case["name"]
        assert abs(r["objective"] - case["objective"]) <= REL_BOUND * abs(case["objective"]), (case["name"], r["objective"], case["objective"])
        assert r["lower_bound"] <= case["objective"] * (1 + 1e-6) + 1e-9      # certified bound never above the optimum
        # returned point is primal feasible for the reference's program (OMC.jl:1554-1561)
        n, k = case["n"], case["k"]
        X, Y, U = r["X"], r["Y"], r["U"]
        assert np.abs(Y - Y.T).max() <= 1e-12
        assert np.linalg.eigvalsh(np.eye(n) - Y).min() >= -1e-6 and np.trace(Y) <= k + 1e-6
        assert np.linalg.eigvalsh(np.block([[Y, U], [U.T, np.eye(k)]])).min() >= -1e-6
        assert np.sqrt((U * U).sum(axis=0)).max() <= 1 + 1e-6                  # OMC.jl:1831-1835 holds although not imposed
        p.close()


def test_relaxation_matches_oracle_iterate_for_iterate(omc):
    """Same algorithm, same iteration count: the GPU trajectory equals the oracle's to rounding."""
    from oracle import relaxation as R
    from oracle.datagen import config_instance
    from oracle.cuts import child_directions
    k, A, mask, g = config_instance("C1", 0)
    p = omc.Problem(k, A, mask, g, "linear")
    rng = np.random.default_rng(5)
    x = rng.standard_normal(10); x /= np.linalg.norm(x); Uh = 0.3 * rng.standard_normal((10, 1))
    cid = p.add_cut(x, Uh)
    for dirs in (["left"], ["right"]):
        for mi in (1, 7, 60):
            r = p.relax_batch([[omc.Cut(cid, x, Uh, dirs)]], omc.default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adapt_every=0, jacobi_tol=1e-13, exact_projection=1))[0]
            ro = R.solve_relaxation(A, mask, g, k, "linear", [(x, Uh, dirs)], opts=R.Options(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adaptive_rho=False))
            assert np.abs(r["X"] - ro["X"]).max() < 1e-9 and np.abs(r["Y"] - ro["Y"]).max() < 1e-9 and np.abs(r["U"] - ro["U"]).max() < 1e-9
            assert abs(r["res_p"] - ro["res_p"]) <= 1e-7 * max(1.0, ro["res_p"]) and abs(r["res_d"] - ro["res_d"]) <= 1e-7 * max(1.0, ro["res_d"])
    p.close()


def test_relaxation_batch_c2_against_oracle_and_properties(omc):
    """BASELINE config 2 (k=1, 50x50): root + both children, compared with the oracle at full size, and the
    domain properties child bound >= parent bound, x -> -x symmetry, batch == singles, warm == cold."""
    from oracle import relaxation as R
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C2", 0)
    p = omc.Problem(k, A, mask, g, "linear", state_pool_capacity=4)
    opts = omc.default_opts()
    root = p.relax_batch([[]], opts, save_ids=[0])[0]
    ro = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8))
    assert abs(root["objective"] - ro["objective"]) <= REL_BOUND * ro["objective"]
    lam, vec, bp, feas = omc.smallest_eigvecs_batch(root["Y"], root["U"], 1)
    assert not feas[0] and lam[0, 0] < -0.5
    x = bp[0]
    c = p.add_cut(x, root["U"]); cneg = p.add_cut(-x, root["U"])
    kids = [[omc.Cut(c, x, root["U"], ["left"])], [omc.Cut(c, x, root["U"], ["right"])]]
    batch = p.relax_batch(kids, opts)
    singles = [p.relax_batch([kid], opts)[0] for kid in kids]
    warm = p.relax_batch(kids, opts, warm_ids=[0, 0])
    flipped = p.relax_batch([[omc.Cut(cneg, -x, root["U"], ["right"])], [omc.Cut(cneg, -x, root["U"], ["left"])]], opts)
    for b in range(2):
        assert batch[b]["termination_status"] == "OPTIMAL"
        assert batch[b]["objective"] >= root["objective"] * (1 - 1e-6)
        assert batch[b]["objective"] == singles[b]["objective"]                       # deterministic
        assert abs(batch[b]["objective"] - warm[b]["objective"]) <= REL_BOUND * batch[b]["objective"]
        assert abs(batch[b]["objective"] - flipped[b]["objective"]) <= REL_BOUND * batch[b]["objective"]
    oc = R.solve_relaxation(A, mask, g, k, "linear", [(x, root["U"], ["right"])], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8))
    assert abs(batch[1]["objective"] - oc["objective"]) <= REL_BOUND * oc["objective"]
    # any feasible rank-k point bounds the relaxation from above
    u, s, vt = np.linalg.svd(np.where(mask, A, 0.0)); Xr = (u[:, :k] * s[:k]) @ vt[:k]
    assert root["objective"] <= p.objective_mse(Xr)[0]
    # state-pool ids of one launch: a record is written by at most one node and not read by another (cross-CTA race otherwise)
    with pytest.raises(RuntimeError, match="appears twice"):
        p.relax_batch(kids, opts, save_ids=[1, 1])
    with pytest.raises(RuntimeError, match="written by another node"):
        p.relax_batch(kids, opts, warm_ids=[0, 2], save_ids=[2, 3])
    p.close()


def test_cutoff_prunes_dominated_node_early(omc):
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C1", 0)
    p = omc.Problem(k, A, mask, g, "linear")
    root = p.relax_batch([[]])[0]
    x = omc.smallest_eigvecs_batch(root["Y"], root["U"], 1)[2][0]
    c = p.add_cut(x, root["U"])
    kid = [[omc.Cut(c, x, root["U"], ["left"])]]
    full = p.relax_batch(kid)[0]
    assert full["objective"] > root["objective"] * 1.05
    cut = p.relax_batch(kid, omc.default_opts(cutoff=root["objective"] * 1.01))[0]
    assert cut["status_code"] == 4 and cut["termination_status"] == "OPTIMAL"
    assert cut["objective"] > root["objective"] * 1.01 and cut["objective"] <= full["objective"] * (1 + 1e-6)
    assert cut["iters"] <= full["iters"]
    p.close()


def test_separation_oracle_matches_dense_eigh(omc):
    from oracle import eigsep as E
    rng = np.random.default_rng(7)
    for n, k in [(10, 1), (30, 2), (50, 1), (100, 3)]:
        B = 5
        Y = np.stack([(lambda M: M @ M.T / n)(rng.standard_normal((n, n))) for _ in range(B)])
        U = rng.standard_normal((B, n, k)) / np.sqrt(n)
        Y[0] = U[0] @ U[0].T + 1e-9 * np.eye(n)               # master-feasible node
        for nev in (1, 2):
            lam, vec, bp, feas = omc.smallest_eigvecs_batch(Y, U, nev)
            for b in range(B):
                lo, vo = E.smallest_eigpairs(Y[b], U[b], nev)
                assert np.abs(lam[b] - lo).max() <= 1e-10
                assert feas[b] == E.master_feasible(Y[b], U[b])
                if b > 0:
                    xo, _ = E.breakpoint_vector(Y[b], U[b], "smallest_1_eigvec" if nev == 1 else "smallest_2_eigvec")
                    assert np.abs(bp[b] - xo).max() <= 1e-8
                    assert abs(np.linalg.norm(vec[b, :, 0]) - 1) < 1e-12


def test_argument_validation_mirrors_reference_errors(omc):
    rng = np.random.default_rng(8)
    A = rng.standard_normal((6, 5)); mask = rng.random((6, 5)) < 0.5
    with pytest.raises(Exception, match="n <= m"):
        omc.Problem(1, A, mask, 80.0)                         # OMC.jl:249-254
    with pytest.raises(ValueError, match="Disjunctive cuts type"):
        omc.Problem(1, A.T, mask.T, 80.0, "linear4")          # OMC.jl:218-224
    with pytest.raises(ValueError, match="Dimension mismatch"):
        omc.Problem(1, A.T, mask, 80.0)                       # OMC.jl:240-246


def test_config4_shape_runs_through_the_l2_resident_block_path(omc):
    """BASELINE config 4 shape (k=3, 100x100, linear3): the (n+m) = 200 PSD block does not fit one SM's shared
    memory and is diagonalised in an L2-resident buffer by the same device code.  Fixed-iteration trajectory vs the
    oracle, then a full solve whose returned point must be feasible and whose certified bound must not exceed it."""
    from oracle import relaxation as R
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C4", 0)
    p = omc.Problem(k, A, mask, g, "linear3")
    rng = np.random.default_rng(11)
    x = rng.standard_normal(100); x /= np.linalg.norm(x); Uh = 0.2 * rng.standard_normal((100, 3))
    cid = p.add_cut(x, Uh)
    dirs = ["inner_left", "right", "left"]
    for cuts_g, cuts_o in (([], []), ([omc.Cut(cid, x, Uh, dirs)], [(x, Uh, dirs)])):
        r = p.relax_batch([cuts_g], omc.default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=20, adapt_every=0, jacobi_tol=1e-13, exact_projection=1),
                          engine="persistent")[0]          # (engine "auto" sends n + m > 104 to the batched engine, tests/test_gpu_big.py)
        ro = R.solve_relaxation(A, mask, g, k, "linear3", cuts_o, opts=R.Options(eps_abs=1e-30, eps_rel=1e-30, max_iter=20, adaptive_rho=False))
        assert np.abs(r["X"] - ro["X"]).max() < 1e-8 and np.abs(r["Y"] - ro["Y"]).max() < 1e-8 and np.abs(r["U"] - ro["U"]).max() < 1e-8
    full = p.relax_batch([[]], omc.default_opts(max_iter=4000), engine="persistent")[0]
    assert full["termination_status"] == "OPTIMAL"
    auto = p.relax_batch([[]], omc.default_opts(max_iter=4000))[0]            # the batched engine agrees with the persistent one
    assert auto["termination_status"] == "OPTIMAL" and abs(auto["objective"] - full["objective"]) <= REL_BOUND * abs(full["objective"])
    X, Y, U = full["X"], full["Y"], full["U"]
    assert np.linalg.eigvalsh(np.eye(100) - Y).min() >= -1e-6 and np.trace(Y) <= k + 1e-6 and np.linalg.eigvalsh(Y).min() >= -1e-6
    assert full["lower_bound"] <= full["objective"] * (1 + 1e-6)
    am = omc.alternating_minimization(p, np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k])
    assert full["objective"] <= p.objective_mse(am["U"] @ am["V"])[0] * (1 + 1e-9)      # bound <= a feasible rank-k value
    lam, vec, bp, feas = omc.smallest_eigvecs_batch(Y, U, 2)                            # config 4 uses smallest_2_eigvec
    assert lam[0, 0] <= lam[0, 1] and abs(np.linalg.norm(vec[0, :, 0]) - 1) < 1e-10
    p.close()


def test_tracked_projection_matches_exact_projection_on_frontier_nodes(omc):
    """The default path tracks the minority spectral side of every PSD block with one block-LOBPCG step per
    iteration (csrc/omc_lowrank.cuh); exact_projection=1 diagonalises every block at every iteration.  On deep
    config-2 frontier nodes both must stop at the same iteration (+1 for the confirming exact iteration) with the
    same bound, the tracked run must do almost all projections on the tracker, and both reproduce the bounds stored in
    tests/golden/c2_frontier.json (which tests/test_frontier_fixture.py pins to the CPU oracle) to the north-star tolerance."""
    import bench, json, os
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C2", 0)
    p = omc.Problem(k, A, mask, g, "linear")
    cuts = bench.load_frontier_fixture(6)
    nodes = [_cuts_for(omc, p, cl) for cl in cuts]
    o_t = omc.default_opts(max_iter=6000)
    o_e = omc.default_opts(max_iter=6000, exact_projection=1)
    f = omc.Frontier(p, nodes); f.relax(o_t); rt = f.fetch(True); prof = f.profile(); f.close()
    re_ = p.relax_batch(nodes, o_e)
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_frontier.json")))["nodes"][:6]
    for a, b, rec in zip(rt, re_, fx):
        assert a["status_code"] == b["status_code"]
        # converged runs agree to 1e-7; a node stopped by max_iter is a point on two slightly different trajectories
        tol = 1e-7 if a["status_code"] == 0 else 1e-5
        assert abs(a["objective"] - b["objective"]) <= tol * abs(b["objective"])
        if rec["gpu"]["status"] == 0 and a["status_code"] == 0:
            # the stored bounds were taken with rho0 = 0.1: two ADMM runs that both stop at eps = 1e-8 agree to a few 1e-7
            assert abs(a["objective"] - rec["gpu"]["objective"]) <= 3 * REL_BOUND * abs(a["objective"])
        assert abs(a["iters"] - b["iters"]) <= max(26, 0.06 * b["iters"])      # one check period, or 6 % on slow nodes
        if a["status_code"] == 0:
            assert np.abs(a["X"] - b["X"]).max() <= 1e-4 and np.abs(a["Y"] - b["Y"]).max() <= 1e-5
        Y, U = a["Y"], a["U"]
        assert np.linalg.eigvalsh(np.eye(50) - Y).min() >= -1e-6 and np.linalg.eigvalsh(Y - U @ U.T).min() >= -1e-6
    assert prof[:, 13].sum() + prof[:, 14].sum() >= 0.98 * (prof[:, 13].sum() + prof[:, 14].sum() + prof[:, 15].sum())   # tracker (incl. idle) did >= 98 % of the projections
    p.close()


def test_alternating_minimization_matches_oracle(omc):
    """K7+K8 against oracle/altmin.py (OMC.jl:1979-2279) on the config 1-3 shapes, from +U0 and -U0 (the reference does
    not normalise the SVD sign, OMC.jl:522-524) and with a cut: same sweep count, same convergence flag, same iterates."""
    from oracle import altmin as AM
    from oracle.datagen import config_instance, CONFIGS
    from oracle.cuts import LABELS
    for cfg in ("C1", "C2", "C3"):
        k, A, mask, g = config_instance(cfg, 0)
        n = A.shape[0]
        ct = CONFIGS[cfg]["cut_type"]
        p = omc.Problem(k, A, mask, g, ct)
        U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
        rng = np.random.default_rng(0)
        x = rng.standard_normal(n); x /= np.linalg.norm(x); Uh = 0.3 * rng.standard_normal((n, k))
        dirs = [LABELS[ct][-1]] * k
        cid = p.add_cut(x, Uh)
        for Ui, gc, oc in ((U0, None, ()), (-U0, None, ()), (U0, [omc.Cut(cid, x, Uh, dirs)], [(x, Uh, dirs)])):
            r = omc.alternating_minimization(p, Ui, gc)
            ro = AM.alternating_minimization(A, n, k, mask, g, True, ct, Ui, oc)
            assert r["converged"] == ro["converged"] and r["n_iters"] == ro["n_iters"], (cfg, r["n_iters"], ro["n_iters"])
            assert np.abs(r["U"] - ro["U"]).max() <= 1e-8 and np.abs(r["V"] - ro["V"]).max() <= 1e-7
            assert np.allclose(r["objectives"], ro["objectives"], rtol=1e-8, atol=0)     # inner U-step ADMM tolerance 1e-9
            assert abs(p.objective_mse(r["U"] @ r["V"])[0] - r["objectives"][-1]) <= 1e-9 * abs(r["objectives"][-1])   # a6 agrees
        p.close()


def test_branch_and_bound_certifies_config2_and_brackets_config1(omc):
    """The host loop (mirror of OMC.jl:700-1073).  Config 2 certifies gap <= 1e-4 in a few nodes: the incumbent's objective
    is evaluate_objective of the returned rank-k X (OMC.jl:2330-2359) and is bracketed by the oracle's root relaxation.
    Config 1 (50 of 100 entries observed) needs thousands of nodes, so it is run for a bounded number of steps under two
    node selections and two frontier batch sizes: every run's lower bound must stay below every run's incumbent, bounds
    must be monotone, and batching must not change the first pops (frontier_batch = 1 is the reference's sequence)."""
    from oracle import relaxation as R
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C2", 0)
    root = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8))
    sol, printlist, inst = omc.matrix_completion_branchandbound(k, A, mask, g, node_selection="bestfirst", disjunctive_cuts_type="linear",
                                                                disjunctive_cuts_breakpoints="smallest_1_eigvec", time_limit=120, verbosity=0)
    assert inst["tree"].now_gap <= 1e-4 + 1e-12
    assert root["objective"] * (1 - 1e-6) <= sol["objective"]
    p = omc.Problem(k, A, mask, g, "linear")
    assert abs(p.objective_mse(sol["X"])[0] - sol["objective"]) <= 1e-9 * abs(sol["objective"])
    assert np.linalg.matrix_rank(sol["X"], tol=1e-8) <= k
    p.close()
    k, A, mask, g = config_instance("C1", 0)
    root = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8))
    runs = []
    for kw in (dict(node_selection="bestfirst", frontier_batch=16), dict(node_selection="breadthfirst", frontier_batch=16),
               dict(node_selection="bestfirst", frontier_batch=1, max_steps=12)):
        kw.setdefault("max_steps", 160)
        sol, printlist, inst = omc.matrix_completion_branchandbound(k, A, mask, g, disjunctive_cuts_type="linear",
                                                                    disjunctive_cuts_breakpoints="smallest_1_eigvec",
                                                                    use_max_steps=True, time_limit=120, verbosity=0, **kw)
        tree = inst["tree"]
        lbs = [row[3] for row in inst["run_log"]]                                          # (explored, counter, remaining, LB, ...)
        assert all(b2 >= b1 - 1e-9 for b1, b2 in zip(lbs, lbs[1:]))                         # OMC.jl:1213-1216: LB never decreases
        runs.append((tree.best_lower_bound, sol["objective"], tree.now_gap))
        assert sol["objective"] >= root["objective"] * (1 - 1e-6)
    los = [r[0] for r in runs if r[0] is not None and np.isfinite(r[0])]
    ups = [r[1] for r in runs]
    assert not los or max(los) <= min(ups) * (1 + 1e-6), runs                               # any lower bound <= any incumbent


def test_config3_shape_rank2_linear2_tracked_vs_oracle(omc):
    """BASELINE config 3 shape without the Shor rows (k = 2, 30 x 30, linear2 cuts: 9 children per split): root and three
    children, default (tracked) path vs the oracle's exact-projection ADMM to the north-star tolerance, and vs the GPU's own
    exact path; child bounds dominate the parent's."""
    from oracle import relaxation as R
    from oracle.datagen import config_instance
    from omc_b200.host import BBNode, create_matrix_cut_child_nodes
    k, A, mask, g = config_instance("C3", 0)
    p = omc.Problem(k, A, mask, g, "linear2")
    root = p.relax_batch([[]])[0]
    ro = R.solve_relaxation(A, mask, g, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8))
    assert root["status_code"] == 0 and abs(root["objective"] - ro["objective"]) <= REL_BOUND * abs(ro["objective"])
    lam, vec, bp, feas = omc.smallest_eigvecs_batch(root["Y"][None], root["U"][None], 1)
    assert not feas[0]
    kids = create_matrix_cut_child_nodes(p, BBNode(node_id=1, parent_id=0, LB=root["objective"], depth=0), bp[0], root["U"], 1, root["objective"])
    assert len(kids) == 9                                                     # P^k children, OMC.jl:2481-2491
    pick = [kids[i] for i in (0, 2, 4, 8)]
    nconv = 0
    rt = p.relax_batch([kd.disjunctive_cuts for kd in pick], omc.default_opts(max_iter=8000))
    re_ = p.relax_batch([kd.disjunctive_cuts for kd in pick], omc.default_opts(max_iter=8000, exact_projection=1))
    for kd, a, b in zip(pick, rt, re_):
        cuts_o = [(c.x, c.Uhat, c.directions) for c in kd.disjunctive_cuts]
        o = R.solve_relaxation(A, mask, g, k, "linear2", cuts_o, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=8000))
        # at the root U = 0, so vhat = 0 and the "middle" region of linear2 is the single point v = 0:
        # that child is degenerate for any first-order method and may stop at max_iter on all three paths alike
        assert a["status_code"] == b["status_code"] == o["status"], (kd.disjunctive_cuts[0].directions, a["status_code"], b["status_code"], o["status"])
        if a["status_code"] == 0:            # iterates cut off by max_iter are not comparable between implementations
            assert abs(a["objective"] - o["objective"]) <= REL_BOUND * abs(o["objective"]), (a["objective"], o["objective"])
            assert abs(a["objective"] - b["objective"]) <= 3e-7 * abs(b["objective"])
            nconv += 1
        assert a["objective"] >= root["objective"] * (1 - 1e-6)
    assert nconv >= 2
    p.close()


def test_alternating_minimization_batch_equals_single_runs(omc):
    """omc_altmin_batch (one CTA per instance: the reference's extra root restarts OMC.jl:529-538, or a popped batch of
    nodes) returns exactly what the single-instance entry returns for every instance, with and without cuts."""
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C3", 0)
    n = A.shape[0]
    p = omc.Problem(k, A, mask, g, "linear2")
    U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
    rng = np.random.default_rng(4)
    x = rng.standard_normal(n); x /= np.linalg.norm(x); Uh = 0.3 * rng.standard_normal((n, k))
    cut = omc.Cut(p.add_cut(x, Uh), x, Uh, ["left", "right"])
    starts = [U0, -U0, U0 + np.abs(U0).max() * rng.standard_normal((n, k)), U0]            # OMC.jl:538 restart
    cutl = [[], [], [], [cut]]
    batch = omc.alternating_minimization_batch(p, starts, cutl)
    for Ui, cl, rb in zip(starts, cutl, batch):
        rs = omc.alternating_minimization(p, Ui, cl)
        assert rb["converged"] == rs["converged"] and rb["n_iters"] == rs["n_iters"]
        assert np.array_equal(rb["U"], rs["U"]) and np.array_equal(rb["V"], rs["V"]) and rb["objectives"] == rs["objectives"]
    p.close()


def test_alternating_minimization_config5_size_properties(omc):
    """BASELINE config 5 size (k = 5, 1000 x 1000, 200 000 observed): the oracle's dense U-step is too slow there, so the
    root alt-min is checked through properties the reference's program implies: objectives[i] is f(U, V) of the returned
    point re-evaluated by the fused reduction (a6), every U-step constraint holds (OMC.jl:2024-2045, 2164-2171), the V-step
    left a stationary V for the U it started from (normal equations OMC.jl:2192-2209 on a sample of columns), and the first
    sweeps decrease the objective."""
    from oracle.datagen import generate_matrix_completion_data
    k, n, m, g = 5, 1000, 1000, 80.0
    A, mask = generate_matrix_completion_data(k, n, m, 200000, 0)
    p = omc.Problem(k, A, mask, g, "linear")
    U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
    r1 = omc.alternating_minimization(p, U0, max_iters=1)
    r3 = omc.alternating_minimization(p, U0, max_iters=3)
    assert r1["n_iters"] == 1 and r3["n_iters"] == 3 and r3["objectives"][0] == r1["objectives"][0]
    assert r3["objectives"][2] <= r3["objectives"][1] <= r3["objectives"][0]
    U, V = r3["U"], r3["V"]
    assert abs(p.objective_mse(U @ V)[0] - r3["objectives"][-1]) <= 1e-9 * abs(r3["objectives"][-1])
    assert U.max() <= 1 + 1e-7 and U.min() >= -1 - 1e-7 and np.sqrt((U * U).sum(axis=0)).max() <= 1 + 1e-6
    for j in range(k):
        assert U[n - k + j:, j].min() >= -1e-7
        for j2 in range(j + 1, k):
            assert np.linalg.norm(U[:, j] + U[:, j2]) <= np.sqrt(2) + 1e-6 and np.linalg.norm(U[:, j] - U[:, j2]) <= np.sqrt(2) + 1e-6
    # V of the first sweep solves (U0_I' U0_I + U0'U0 / gamma) v_j = U0_I' A_I,j   (checked on 25 columns)
    V1 = r1["V"]
    G0 = U0.T @ U0 / g
    for j in range(0, m, 40):
        I = mask[:, j]
        lhs = U0[I].T @ U0[I] + G0
        assert np.abs(lhs @ V1[:, j] - U0[I].T @ A[I, j]).max() <= 1e-9 * max(1.0, np.abs(A[I, j]).max())
    p.close()


def test_shor_minor_indexes_bit_exact(omc):
    """Row a10: the Shor minor index list and the SOC coordinate list (OMC.jl:2545-2612, 648-665) from the GPU equal the
    oracle's restatement element for element -- same tuples in the same order -- on random masks (empty, full, ragged), on
    every ordering of the num_entries_present list, and on the config 3 mask (180 k minors)."""
    from oracle import shor as S
    from oracle.datagen import config_instance
    rng = np.random.default_rng(7)
    cases = [(2, 2, 1.0), (3, 5, 0.0), (4, 7, 0.5), (7, 9, 0.3), (9, 13, 0.8), (12, 70, 0.5), (6, 6, 1.0)]
    lists = ([4], [3], [2], [1], [0], [1, 2, 3, 4], [4, 3, 2, 1, 0], [2, 4, 2])
    for n, m, dens in cases:
        mask = rng.random((n, m)) < dens
        p = omc.Problem(1, rng.standard_normal((n, m)), mask, 80.0)
        for pl in lists:
            tg, sg = omc.shor_constraint_indexes(p, pl)
            want = np.array(S.shor_constraint_indexes(mask, pl), dtype=np.int64).reshape(-1, 4) - 1
            assert tg.shape == want.shape and np.array_equal(tg, want), (n, m, dens, pl)
            cov = np.zeros((n, m), bool)
            for i1, i2, j1, j2 in want:
                cov[i1, j1] = cov[i1, j2] = cov[i2, j1] = cov[i2, j2] = True
            soc_want = np.array([(i, j) for j in range(m) for i in range(n) if not cov[i, j]], dtype=np.int64).reshape(-1, 2)
            assert np.array_equal(sg, soc_want), (n, m, dens, pl)
        p.close()
    k, A, mask, g = config_instance("C3", 0)
    p = omc.Problem(k, A, mask, g, "linear2")
    tg, sg = omc.shor_constraint_indexes(p, [1, 2, 3, 4])
    want = np.array(S.shor_constraint_indexes(mask, [1, 2, 3, 4]), dtype=np.int64).reshape(-1, 4) - 1
    assert len(want) > 100000 and np.array_equal(tg, want)
    p.close()


def test_edge_shapes_smallest_full_mask_and_k_equal_n(omc):
    """Smallest sizes, fully observed masks, k = n and k = 2 with tall-thin A: bound vs oracle, 1e-6 relative."""
    from oracle import relaxation as R
    from oracle.datagen import generate_matrix_completion_data
    for (n, m, k, nobs) in [(2, 2, 1, 4), (2, 3, 2, 5), (3, 7, 2, 10), (5, 5, 1, 25), (4, 9, 4, 20), (3, 3, 1, 7), (4, 4, 1, 12)]:
        A, mask = generate_matrix_completion_data(k, n, m, nobs, 3)
        p = omc.Problem(k, A, mask, 20.0, "linear")
        r = p.relax_batch([[]], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=200000))[0]
        ro = R.solve_relaxation(A, mask, 20.0, k, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=200000))
        assert r["termination_status"] == "OPTIMAL", (n, m, k)
        assert abs(r["objective"] - ro["objective"]) <= REL_BOUND * abs(ro["objective"]) + 1e-9, (n, m, k, r["objective"], ro["objective"])
        assert r["lower_bound"] <= ro["objective"] * (1 + 1e-6) + 1e-9
        p.close()


def test_ragged_frontier_zero_to_capacity_cuts(omc):
    """One batch holding nodes with 0, 1, 5 and 64 (the capacity) cuts: batch == singles, the 5-cut node equals the
    oracle, bounds grow along a nested cut chain, and capacity + 1 cuts is refused."""
    from oracle import relaxation as R
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C1", 0)
    n = A.shape[0]
    p = omc.Problem(k, A, mask, g, "linear")
    rng = np.random.default_rng(17)
    ustar = rng.standard_normal(n); ustar /= np.linalg.norm(ustar)
    chain, ocuts = [], []
    for _ in range(65):
        x = rng.standard_normal(n); x /= np.linalg.norm(x)
        Uh = rng.uniform(-0.5, 0.5) * x[:, None]                   # vhat = Uh' x in (-0.5, 0.5)
        d = ["left"] if float(x @ ustar) <= float(Uh[:, 0] @ x) else ["right"]   # u* stays feasible down the chain
        chain.append(omc.Cut(p.add_cut(x, Uh), x, Uh, d)); ocuts.append((x, Uh, d))
    opts = omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=200000)
    nodes = [chain[:0], chain[:1], chain[:5], chain[:64]]
    batch = p.relax_batch(nodes, opts)
    for b, nd in enumerate(nodes):
        assert batch[b]["termination_status"] == "OPTIMAL", b
        assert batch[b]["objective"] == p.relax_batch([nd], opts)[0]["objective"]
        if b:
            assert batch[b]["objective"] >= batch[b - 1]["objective"] * (1 - 1e-6)
    ro = R.solve_relaxation(A, mask, g, k, "linear", ocuts[:5], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=200000))
    assert abs(batch[2]["objective"] - ro["objective"]) <= REL_BOUND * ro["objective"]
    # the rank-1 point built from u* is feasible for every node: it bounds all of them from above
    uu = ustar[:, None]
    Xs = uu @ (uu.T @ np.where(mask, A, 0.0))
    assert batch[3]["objective"] <= p.objective_mse(Xs)[0] * (1 + 1e-9)
    # capacity + 1 cuts: the persistent engine refuses, engine "auto" hands the node to the batched engine (no cut cap)
    with pytest.raises(Exception, match="exceeds the 64 the persistent engine supports"):
        p.relax_batch([chain[:65]], opts, engine="persistent")
    deep = p.relax_batch([chain[:65]], opts)[0]
    assert deep["termination_status"] == "OPTIMAL" and deep["objective"] >= batch[3]["objective"] * (1 - 1e-6)
    assert deep["objective"] <= p.objective_mse(Xs)[0] * (1 + 1e-9)
    ro65 = R.solve_relaxation(A, mask, g, k, "linear", ocuts[:65], opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=200000))
    assert abs(deep["objective"] - ro65["objective"]) <= REL_BOUND * ro65["objective"]
    p.close()


def test_tracked_equals_exact_projection_over_a_sweep_of_small_shapes(omc):
    """61 shapes from 2 x 2 to 24 x 72, k = 1..4: the default (tracked) path and the exact-projection path both reach
    OPTIMAL with bounds within 1e-6 relative and comparable iteration counts.  (This sweep found three defects that the
    BASELINE shapes never touch: per-index scratch sized by an 8-row block, the cut Gram staging of a small problem with
    more than sqrt(buf0) cuts, and cold eigensolves run at the loose warm-start tolerance.)"""
    import itertools
    from oracle.datagen import generate_matrix_completion_data
    shapes = [(2, 2, 1), (2, 3, 2), (3, 7, 2), (5, 5, 1), (4, 9, 4), (3, 3, 1), (4, 4, 1)]
    shapes += [(n, n * mm, k) for n, mm, k in itertools.product((6, 8, 10, 12, 16, 24), (1, 2, 3), (1, 2, 3))]
    for (n, m, k) in shapes:
        A, mask = generate_matrix_completion_data(k, n, m, min(max(n + m, int(0.6 * n * m)), n * m), 3)
        p = omc.Problem(k, A, mask, 20.0, "linear")
        rt, re_ = [p.relax_batch([[]], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=30000, exact_projection=ex))[0] for ex in (0, 1)]
        assert rt["termination_status"] == re_["termination_status"] == "OPTIMAL", (n, m, k)
        assert abs(rt["objective"] - re_["objective"]) <= REL_BOUND * abs(re_["objective"]), (n, m, k, rt["objective"], re_["objective"])
        assert rt["iters"] <= 1.5 * re_["iters"] + 50, (n, m, k, rt["iters"], re_["iters"])
        p.close()


def test_cut_chains_of_every_type_and_rank_against_the_oracle(omc):
    """k = 1..3, all three disjunctive cut types (OMC.jl:1580-1683), chains of 3 and 10 cuts: bound of the tracked and
    of the exact path vs the oracle, 1e-6 relative.  Chains on which the oracle itself does not reach OPTIMAL (linear3 /
    right carries the reference's quirk Q1 and can make a node infeasible) are skipped."""
    import itertools
    from oracle import relaxation as R
    from oracle.datagen import generate_matrix_completion_data
    checked = 0
    for (n, m, k), ct, L in itertools.product([(4, 4, 1), (6, 9, 2), (8, 8, 3)], ("linear", "linear2", "linear3"), (3, 10)):
        rng = np.random.default_rng(100 * n + 10 * k + L)
        A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
        cuts = _feasible_chain(ct, n, k, L, rng)
        ro = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000))
        if ro["status"] != R.STATUS_OPTIMAL:
            continue
        p = omc.Problem(k, A, mask, 20.0, ct)
        gc = [omc.Cut(p.add_cut(x, Uh), x, Uh, d) for x, Uh, d in cuts]
        for ex in (0, 1):
            r = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, exact_projection=ex))[0]
            assert r["termination_status"] == "OPTIMAL", (n, m, k, ct, L, ex)
            assert abs(r["objective"] - ro["objective"]) <= REL_BOUND * abs(ro["objective"]), (n, m, k, ct, L, ex, r["objective"], ro["objective"])
        p.close()
        checked += 1
    assert checked >= 15


def test_infeasible_node_is_certified_like_the_oracle(omc):
    """A 10-cut linear3 chain that the reference's `right` quirk (OMC.jl:1675, Q1) makes infeasible.  Default build
    (infeasibility by bound): the certified lower bound exceeds 1/2 ||P(A)||^2, which no feasible node can, and
    the node comes back INFEASIBLE (-> feasible = false, OMC.jl:1921-1935) at the oracle's iteration (same rule,
    Options.infeasible_by_bound) on the exact path and near it on the tracked path.  With the opt-in d mu certificate the
    node stops where the oracle's certificate does.  With `fix_linear3_right` the same chain is feasible again."""
    from oracle import relaxation as R
    from oracle.datagen import generate_matrix_completion_data
    flags = omc.build_flags()
    assert flags & 3, "the shipped build must be able to report MOI.INFEASIBLE"
    n, m, k, ct, L = 6, 9, 2, "linear3", 10
    rng = np.random.default_rng(100 * n + 10 * k + L)
    A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
    cuts = _feasible_chain(ct, n, k, L, rng)
    ro = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000,
                                                                         infeasible_by_bound=bool(flags & 2)))
    assert ro["status"] == R.STATUS_INFEASIBLE and not ro["feasible"]
    p = omc.Problem(k, A, mask, 20.0, ct)
    gc = [omc.Cut(p.add_cut(x, Uh), x, Uh, d) for x, Uh, d in cuts]
    ex = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, exact_projection=1))[0]
    tr = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000))[0]
    assert ex["termination_status"] == "INFEASIBLE" and not ex["feasible"] and abs(ex["iters"] - ro["iters"]) <= 0.1 * ro["iters"] + 25, (ex["iters"], ro["iters"])
    assert tr["termination_status"] == "INFEASIBLE" and not tr["feasible"] and tr["iters"] <= 2 * ro["iters"] + 50, (tr["iters"], ro["iters"])
    fixed = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, fix_linear3_right=1))[0]
    rf = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, fix_linear3_right=True,
                                                                         infeasible_by_bound=bool(flags & 2)))
    assert rf["status"] == R.STATUS_OPTIMAL and fixed["termination_status"] == "OPTIMAL"
    assert abs(fixed["objective"] - rf["objective"]) <= REL_BOUND * rf["objective"]
    p.close()
