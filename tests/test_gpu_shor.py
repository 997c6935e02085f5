"""GPU parity tests of the Shor valid-inequality rows (K4: OMC.jl:1503-1552, 1755-1846) in the batched engine, through the
C ABI (omc_problem_set_shor / omc_frontier_fetch_shor), against oracle/shor_relax.py (same program, exact projections).

Tolerance: 1e-5 relative on the bound at eps 1e-7 (the program is a linear objective over many small cones: the optimal
value is flatter in the iterate than the plain relaxation's; observed 2e-7 .. 5e-6), 1e-4 absolute on X, 2e-3 on W."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def omc():
    import omc_b200
    omc_b200.init(0)
    return omc_b200


def _structure(mask, pattern):
    from oracle import shor as SI
    n, m = mask.shape
    minors = [tuple(v - 1 for v in t) for t in SI.shor_constraint_indexes(mask, pattern)] if pattern else []
    cov = np.zeros((n, m), bool)
    for (i1, i2, j1, j2) in minors:
        cov[i1, j1] = cov[i1, j2] = cov[i2, j1] = cov[i2, j2] = True
    return minors, [(i, j) for i in range(n) for j in range(m) if not cov[i, j]]


@pytest.mark.parametrize("k,n,m,nidx,seed,pattern", [(1, 5, 6, 18, 3, [1, 2, 3, 4]), (2, 5, 6, 20, 3, [1, 2, 3, 4]),
                                                      (1, 6, 7, 22, 2, []), (2, 6, 7, 26, 1, [4]), (3, 5, 6, 22, 2, [3, 4])])
def test_shor_rows_match_oracle(omc, k, n, m, nidx, seed, pattern):
    from oracle import relaxation as R, shor_relax as SR
    from oracle.datagen import generate_matrix_completion_data
    A, mask = generate_matrix_completion_data(k, n, m, nidx, seed)
    minors, soc = _structure(mask, pattern)
    ref = SR.solve_relaxation_shor(A, mask, 20.0, k, minors, soc, opts=R.Options(eps_abs=1e-7, eps_rel=1e-7, max_iter=60000))
    assert ref["status"] == R.STATUS_OPTIMAL
    p = omc.Problem(k, A, mask, 20.0)
    p.set_shor(minors, soc)
    fr = p.frontier([[]])
    assert fr.stats()["engine"] == "batched"
    fr.relax(omc.default_opts(eps_abs=1e-7, eps_rel=1e-7, max_iter=60000))
    r = fr.fetch()[0]
    W, Xt = fr.fetch_shor()
    assert r["termination_status"] == "OPTIMAL", r
    assert abs(r["objective"] - ref["objective"]) <= 1e-5 * abs(ref["objective"]), (r["objective"], ref["objective"])
    assert r["iters"] <= 1.3 * ref["iters"] + 100, (r["iters"], ref["iters"])
    # (W is not unique where it costs nothing -- unobserved coordinates of a column whose Theta_jj has slack -- hence the looser pin)
    assert np.abs(r["X"] - ref["X"]).max() <= 1e-4 and np.abs(W[0] - ref["W"]).max() <= 2e-3
    assert np.abs(Xt[0].sum(axis=0) - r["X"]).max() <= 1e-6           # X = sum_t Xt (OMC.jl:1492-1493)
    assert r["lower_bound"] == float("-inf")                          # no certified bound with these rows (include/omc_b200.h)
    # Theta_jj = sum_i W_ij and W >= X^2 hold at the returned point (OMC.jl:1757-1767): the objective is then >= the plain one
    assert (W[0] >= r["X"] ** 2 - 1e-5).all()
    fr.close()
    # removing the rows gives the plain relaxation back
    p.set_shor([], [])
    r0 = p.relax_batch([[]], omc.default_opts(eps_abs=1e-7, eps_rel=1e-7, max_iter=60000), engine="batched")[0]
    ro = R.solve_relaxation(A, mask, 20.0, k, opts=R.Options(eps_abs=1e-7, eps_rel=1e-7, max_iter=60000))
    assert abs(r0["objective"] - ro["objective"]) <= 1e-6 * abs(ro["objective"])
    assert r["objective"] >= r0["objective"] * (1 - 1e-5)             # the inequalities tighten


def test_shor_rows_with_cut_nodes_batch_equals_singles(omc):
    """Two cut nodes (chains of 4 and 3 linear2 cuts) relaxed in one batch == relaxed alone (bit for bit) == the oracle."""
    from conftest import feasible_chain
    from oracle import relaxation as R, shor_relax as SR
    from oracle.datagen import generate_matrix_completion_data
    k, n, m = 2, 6, 7
    A, mask = generate_matrix_completion_data(k, n, m, 26, 1)
    minors, soc = _structure(mask, [3, 4])
    nodes = [feasible_chain("linear2", n, k, 4, np.random.default_rng(2)), feasible_chain("linear2", n, k, 3, np.random.default_rng(3))]
    p = omc.Problem(k, A, mask, 20.0, "linear2")
    p.set_shor(minors, soc)
    gnodes = [[omc.Cut(p.add_cut(x, Uh), x, Uh, d) for x, Uh, d in nd] for nd in nodes]
    o = omc.default_opts(eps_abs=1e-7, eps_rel=1e-7, max_iter=60000)
    both = p.relax_batch(gnodes, o)
    for q, nd in enumerate(nodes):
        one = p.relax_batch([gnodes[q]], o)[0]
        assert one["objective"] == both[q]["objective"] and one["iters"] == both[q]["iters"]
        ref = SR.solve_relaxation_shor(A, mask, 20.0, k, minors, soc, "linear2", nd, opts=R.Options(eps_abs=1e-7, eps_rel=1e-7, max_iter=60000))
        assert ref["status"] == R.STATUS_OPTIMAL and both[q]["termination_status"] == "OPTIMAL"
        assert abs(both[q]["objective"] - ref["objective"]) <= 1e-5 * abs(ref["objective"]), (q, both[q]["objective"], ref["objective"])


def test_shor_argument_checks(omc):
    from oracle.datagen import generate_matrix_completion_data
    A, mask = generate_matrix_completion_data(1, 5, 6, 18, 3)
    p = omc.Problem(1, A, mask, 20.0)
    with pytest.raises(RuntimeError, match="out of range"):
        p.set_shor([(0, 7, 0, 1)], [])
    with pytest.raises(RuntimeError, match="covered"):
        p.set_shor([(0, 1, 0, 1)], [(0, 0)])
    p.set_shor([(0, 1, 0, 1)], [(2, 2)])
    with pytest.raises(RuntimeError, match="batched engine only"):
        p.frontier([[]], engine="persistent")


def test_branchandbound_with_shor_rows_certifies_same_optimum(omc):
    """matrix_completion_branchandbound(...; add_Shor_valid_inequalities = true) on a test-scale instance reaches the same
    certified optimum as without the rows (they are valid for every rank-k point), with a root bound at least as tight."""
    from oracle.datagen import generate_matrix_completion_data
    k, n, m = 1, 6, 7
    A, mask = generate_matrix_completion_data(k, n, m, 26, 1)
    kw = dict(node_selection="bestfirst", disjunctive_cuts_type="linear", disjunctive_cuts_breakpoints="smallest_1_eigvec",
              gap=1e-3, max_steps=400, use_max_steps=True, time_limit=300,
              relax_opts=omc.default_opts(eps_abs=1e-7, eps_rel=1e-7, max_iter=20000))
    s0, _, i0 = omc.matrix_completion_branchandbound(k, A, mask, 20.0, **kw)
    s1, _, i1 = omc.matrix_completion_branchandbound(k, A, mask, 20.0, add_Shor_valid_inequalities=True, **kw)
    assert abs(s1["objective"] - s0["objective"]) <= 2e-3 * abs(s0["objective"])
    assert i1["run_log"][0][3] >= i0["run_log"][0][3] * (1 - 1e-5)           # root lower bound
    assert i1["run_details"]["nodes_explored"] <= i0["run_details"]["nodes_explored"] + 2
    assert len(i1["Shor_info"]["constraints_indexes"]) > 0


def test_shor_rows_on_the_config3_mask_properties(omc):
    """BASELINE config 3 (k = 2, 30 x 30, pattern [1, 2, 3, 4]: 177 853 minors, 355 706 moment blocks) is out of the exact oracle's
    reach, so the full size is pinned by size-independent properties: the root converges; X = sum_t Xt; W >= X^2 everywhere
    (implied by the (k+1) blocks, OMC.jl:1810-1826); plain bound <= Shor bound <= objective of a rank-k point (validity)."""
    from oracle.datagen import config_instance
    k, A, mask, g = config_instance("C3", 0)
    p = omc.Problem(k, A, mask, g, "linear2")
    o = omc.default_opts(eps_abs=1e-6, eps_rel=1e-6, max_iter=40000)
    plain = p.relax_batch([[]], o, engine="batched")[0]
    minors, soc = omc.shor_constraint_indexes(p, [1, 2, 3, 4])
    assert len(minors) > 170000 and len(soc) == 0
    p.set_shor(minors, soc)
    fr = p.frontier([[]])
    fr.relax(o)
    r = fr.fetch()[0]
    W, Xt = fr.fetch_shor()
    assert r["termination_status"] == "OPTIMAL", r["iters"]
    assert np.abs(Xt[0].sum(axis=0) - r["X"]).max() <= 1e-6
    # (at eps 1e-6 the cone rows hold for the projections, the iterate is within the primal residual of them: W >= X^2 is
    # tight at most coordinates of the optimum, so the check carries a 2 % slack)
    assert (W[0] >= r["X"] ** 2 - 0.02 * (1.0 + r["X"] ** 2)).all(), float((r["X"] ** 2 - W[0]).max())
    U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
    am = omc.alternating_minimization(p, U0)
    ub = p.objective_mse(am["U"] @ am["V"])[0]
    assert plain["objective"] * (1 - 1e-4) <= r["objective"] <= ub * (1 + 1e-6), (plain["objective"], r["objective"], ub)


def test_violated_shor_minors_match_the_reference_expression(omc):
    """generate_violated_Shor_minors (OMC.jl:2614-2640) on the GPU vs the loop restatement: same tuples in the same order, scores
    bit for bit (the kernel rounds the two products and their difference separately, like the reference's expression), at a
    small shape and at the config-3 size (177 853 candidates)."""
    from oracle import shor as SI
    from oracle.datagen import generate_matrix_completion_data, config_instance
    rng = np.random.default_rng(3)
    for (k, n, m, nidx, seed, n_minors, n_ex) in ((2, 8, 9, 40, 1, 25, 40), (1, 6, 7, 22, 2, 1000, 0)):
        A, mask = generate_matrix_completion_data(k, n, m, nidx, seed)
        p = omc.Problem(k, A, mask, 20.0)
        cand, _ = omc.shor_constraint_indexes(p, [1, 2, 3, 4])
        Xt = rng.standard_normal((k, n, m))
        ex = cand[rng.choice(len(cand), size=n_ex, replace=False)] if n_ex else np.zeros((0, 4), np.int32)
        sc, tp = omc.generate_violated_Shor_minors(p, Xt, cand, ex, n_minors)
        ref = SI.violated_minors(Xt, cand, ex, n_minors)
        assert len(sc) == len(ref) == min(n_minors, len(cand) - n_ex)
        assert [tuple(int(v) for v in t) for t in tp] == [t for _, t in ref]
        assert np.array_equal(sc, np.array([s for s, _ in ref]))
        p.close()
    k, A, mask, g = config_instance("C3", 0)
    p = omc.Problem(k, A, mask, g, "linear2")
    cand, _ = omc.shor_constraint_indexes(p, [1, 2, 3, 4])
    Xt = rng.standard_normal((k, 30, 30))
    ex = cand[rng.choice(len(cand), size=1000, replace=False)]
    sc, tp = omc.generate_violated_Shor_minors(p, Xt, cand, ex, 100)
    d = np.abs(Xt[:, cand[:, 0], cand[:, 2]] * Xt[:, cand[:, 1], cand[:, 3]] - Xt[:, cand[:, 0], cand[:, 3]] * Xt[:, cand[:, 1], cand[:, 2]])
    score = d[0] + d[1] if k == 2 else d.sum(axis=0)
    exs = {tuple(t) for t in ex.tolist()}
    keep = np.array([tuple(t) not in exs for t in cand.tolist()])
    order = sorted(np.flatnonzero(keep).tolist(), key=lambda q: (score[q], tuple(cand[q].tolist())), reverse=True)[:100]
    assert np.array_equal(tp, cand[order]) and np.array_equal(sc, score[order])


def test_branchandbound_with_iterative_shor_rows(omc):
    """add_Shor_valid_inequalities_iterative = true (OMC.jl:670-675, 956-982, 2495-2540): the root carries no minors and an RSOC
    row on every coordinate; splits add the most violated minors to the children's shared list (generate_violated_Shor_minors on
    the GPU); nodes are relaxed in groups that share a row structure.  A short run: bounds stay ordered, the lists stay
    consistent (no duplicate minor, RSOC coordinates = the uncovered ones), the incumbent is the one the plain run finds."""
    from oracle.datagen import generate_matrix_completion_data
    k, n, m = 1, 8, 9
    A, mask = generate_matrix_completion_data(k, n, m, 30, 2)
    kw = dict(node_selection="bestfirst", disjunctive_cuts_type="linear", disjunctive_cuts_breakpoints="smallest_1_eigvec",
              gap=1e-4, max_steps=14, use_max_steps=True, time_limit=120, frontier_batch=4,
              relax_opts=omc.default_opts(eps_abs=1e-6, eps_rel=1e-6, max_iter=20000))
    s0, _, i0 = omc.matrix_completion_branchandbound(k, A, mask, 20.0, **kw)
    s1, _, i1 = omc.matrix_completion_branchandbound(k, A, mask, 20.0, add_Shor_valid_inequalities=True,
                                                     add_Shor_valid_inequalities_iterative=True, update_Shor_indices_n_minors=20, **kw)
    assert abs(s1["objective"] - s0["objective"]) <= 1e-3 * abs(s0["objective"])
    assert i1["tree"].best_lower_bound <= i1["tree"].best_upper_bound * (1 + 1e-6)
    assert i1["tree"].best_lower_bound >= i0["run_log"][0][3] * (1 - 1e-4)        # at least the plain root bound
    assert len(i1["Shor_info"]["constraints_indexes"]) == 0 and len(i1["Shor_info"]["SOC_constraints_indexes"]) == n * m
    assert i1["run_details"]["add_Shor_valid_inequalities_iterative"] is True
    assert i1["run_details"]["nodes_relax_feasible_split"] >= 1 and i1["run_details"].get("Shor_indices_updates", 0) >= 1
    open_infos = {id(nd.Shor_info): nd.Shor_info for nd in i1["open_nodes"]}
    assert len(open_infos) >= 1
    for info in open_infos.values():
        mn, sc = info.constraints_indexes, info.SOC_constraints_indexes
        cov = np.zeros((n, m), bool)
        for (a1, a2, b1, b2) in mn:
            cov[a1, b1] = cov[a1, b2] = cov[a2, b1] = cov[a2, b2] = True
        assert len(mn) % 20 == 0 and len(np.unique(mn, axis=0)) == len(mn)
        assert not cov[sc[:, 0], sc[:, 1]].any() and cov.sum() + len(sc) == n * m
