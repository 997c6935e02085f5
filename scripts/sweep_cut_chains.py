"""Robustness sweep over the entry points next to the relaxation kernel, at small and odd shapes:
 A. relaxation with cut chains of every cut type (feasible by construction) vs the oracle, tracked and exact;
 B. warm vs cold starts on small shapes;  C. separation oracle / PSD projection at tiny n;  D. alt-min at tiny shapes.
Run with --cpu to time only the oracle side (no GPU)."""
import sys, os, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import relaxation as R, eigsep as E, altmin as AM
from oracle.datagen import generate_matrix_completion_data
from oracle.cuts import LABELS

CPU = "--cpu" in sys.argv
if not CPU:
    import omc_b200 as omc
    omc.init(0)


def region(ct, v, h):
    a = abs(h)
    if ct == "linear":
        return "left" if v <= h else "right"
    if ct == "linear2":
        return "left" if v <= -a else ("middle" if v <= a else "right")
    return "left" if v <= -a else ("inner_left" if v <= 0 else ("inner_right" if v <= a else "right"))


def chain(ct, n, k, L, rng):
    # hidden feasible factor: orthonormal columns whose bottom k x k block is a positive diagonal (the sign normalisation
    # U[i, j] >= 0 for i >= n - k + j of the relaxation, OMC.jl:1442-1449)
    W, _ = np.linalg.qr(rng.standard_normal((n - k, k)))
    th = rng.uniform(0.3, 1.2, size=k)
    Us = np.vstack([W * np.cos(th), np.diag(np.sin(th))])
    out = []
    for _ in range(L):
        x = rng.standard_normal(n); x /= np.linalg.norm(x)
        Uh = rng.uniform(-0.4, 0.4, size=(1, k)) * x[:, None]     # vhat_j = Uh[:, j]' x
        out.append((x, Uh, [region(ct, float(x @ Us[:, j]), float(Uh[:, j] @ x)) for j in range(k)]))
    return out


bad = 0
t0 = time.time()
# ---- A
for (n, m, k), ct, L in itertools.product([(4, 4, 1), (6, 9, 2), (8, 8, 3), (10, 20, 2)], ("linear", "linear2", "linear3"), (1, 3, 10)):
    rng = np.random.default_rng(100 * n + 10 * k + L)
    A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 5)
    cuts = chain(ct, n, k, L, rng)
    ro = R.solve_relaxation(A, mask, 20.0, k, ct, cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000))
    line = [(n, m, k), ct, L, "oracle", ro["status"], ro["iters"], "%.9g" % ro["objective"]]
    ok = True
    ref_ok = ro["status"] == 0     # linear3 / right carries the reference's quirk Q1: such chains may be infeasible or stall
    if not CPU:
        p = omc.Problem(k, A, mask, 20.0, ct)
        gc = [omc.Cut(p.add_cut(x, Uh), x, Uh, d) for x, Uh, d in cuts]
        for ex in (0, 1):
            r = p.relax_batch([gc], omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, exact_projection=ex))[0]
            if ref_ok:
                ok = ok and r["termination_status"] == "OPTIMAL" and abs(r["objective"] - ro["objective"]) <= 1e-6 * abs(ro["objective"])
            else:
                ok = ok and r["status_code"] == ro["status"]
            line += ["exact" if ex else "tracked", r["termination_status"], r["iters"], "%.9g" % r["objective"]]
        p.close()
    bad += not ok
    print("A", ("OK " if ok else "BAD") + ("" if ref_ok else " (oracle not optimal)"), *line, flush=True)
print("A done %.1fs" % (time.time() - t0), flush=True)

if not CPU and "--only-a" not in sys.argv:
    # ---- B: warm vs cold
    for (n, m, k) in [(3, 3, 1), (4, 6, 2), (6, 6, 1), (8, 12, 2), (12, 12, 3)]:
        A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.6 * n * m)), 7)
        p = omc.Problem(k, A, mask, 20.0, "linear", state_pool_capacity=2)
        o = omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=60000)
        root = p.relax_batch([[]], o, save_ids=[0])[0]
        lam, vec, bp, feas = omc.smallest_eigvecs_batch(root["Y"], root["U"], 1)
        x = bp[0]
        cid = p.add_cut(x, root["U"])
        kids = [[omc.Cut(cid, x, root["U"], list(d))] for d in itertools.product(("left", "right"), repeat=k)]
        cold = p.relax_batch(kids, o)
        warm = p.relax_batch(kids, o, warm_ids=[0] * len(kids))
        for c_, w_ in zip(cold, warm):
            ok = c_["termination_status"] == w_["termination_status"] == "OPTIMAL" and abs(c_["objective"] - w_["objective"]) <= 1e-6 * abs(c_["objective"]) \
                and c_["objective"] >= root["objective"] * (1 - 1e-6)
            bad += not ok
            print("B", "OK " if ok else "BAD", (n, m, k), "feas", bool(feas[0]), "root %.9g" % root["objective"], "cold", c_["termination_status"], c_["iters"], "%.9g" % c_["objective"],
                  "warm", w_["termination_status"], w_["iters"], "%.9g" % w_["objective"], flush=True)
        p.close()
    # ---- C: separation oracle and PSD projection at tiny sizes
    rng = np.random.default_rng(3)
    for n, k in [(2, 1), (2, 2), (3, 1), (3, 2), (5, 3), (7, 1), (9, 4), (17, 2)]:
        B = 4
        Y = np.stack([(lambda M: M @ M.T / n)(rng.standard_normal((n, n))) for _ in range(B)])
        U = rng.standard_normal((B, n, k)) / np.sqrt(n)
        Y[0] = U[0] @ U[0].T + 1e-9 * np.eye(n)
        for nev in (1, 2):
            if nev > n:
                continue
            lam, vec, bp, feas = omc.smallest_eigvecs_batch(Y, U, nev)
            ok = True
            for b in range(B):
                lo, vo = E.smallest_eigpairs(Y[b], U[b], nev)
                ok = ok and np.abs(lam[b] - lo).max() <= 1e-10 and feas[b] == E.master_feasible(Y[b], U[b])
                if b > 0:
                    xo, _ = E.breakpoint_vector(Y[b], U[b], "smallest_1_eigvec" if nev == 1 else "smallest_2_eigvec")
                    ok = ok and np.abs(bp[b] - xo).max() <= 1e-8
            bad += not ok
            print("C", "OK " if ok else "BAD", "eigsep", (n, k, nev), flush=True)
    for N in (1, 2, 3, 5, 7, 8, 9, 15, 16, 17):
        V = rng.standard_normal((3, N, N)); V = V + np.transpose(V, (0, 2, 1))
        P, lam, sw, _ = omc.psd_project_batch(V)
        ok = all(np.abs(P[b] - R.psd_project(V[b])).max() <= 1e-12 * max(1.0, np.abs(V[b]).max()) for b in range(3))
        bad += not ok
        print("C", "OK " if ok else "BAD", "psd", N, flush=True)
    # ---- D: alt-min at tiny shapes
    for (n, m, k), ct in itertools.product([(2, 2, 1), (3, 5, 2), (4, 4, 3), (6, 9, 2)], ("linear", "linear3")):
        A, mask = generate_matrix_completion_data(k, n, m, max(n + m, int(0.7 * n * m)), 9)
        p = omc.Problem(k, A, mask, 20.0, ct)
        U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
        rng = np.random.default_rng(1)
        x = rng.standard_normal(n); x /= np.linalg.norm(x); Uh = 0.3 * rng.standard_normal((n, k))
        dirs = [LABELS[ct][-1]] * k
        cid = p.add_cut(x, Uh)
        for gc, oc in ((None, ()), ([omc.Cut(cid, x, Uh, dirs)], [(x, Uh, dirs)])):
            r = omc.alternating_minimization(p, U0, gc)
            ro = AM.alternating_minimization(A, n, k, mask, 20.0, True, ct, U0, oc)
            ok = r["converged"] == ro["converged"] and r["n_iters"] == ro["n_iters"] and np.abs(r["U"] - ro["U"]).max() <= 1e-7 \
                and np.allclose(r["objectives"], ro["objectives"], rtol=1e-7, atol=0)
            bad += not ok
            print("D", "OK " if ok else "BAD", (n, m, k), ct, "cut" if gc else "nocut", r["n_iters"], ro["n_iters"], r["converged"], ro["converged"],
                  "%.9g %.9g" % (r["objectives"][-1], ro["objectives"][-1]), flush=True)
        p.close()
print("anomalies:", bad, "total %.1fs" % (time.time() - t0))
