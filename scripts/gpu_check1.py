"""First GPU bring-up: peaks, eigensolver self-test, mask/objective parity, relaxation parity vs oracle."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from omc_b200 import Problem, Cut, default_opts
from oracle.datagen import config_instance, CONFIGS
from oracle import relaxation as R, objective as O, mask as M, eigsep as E
from oracle.cuts import child_directions

omc_b200.init(0)
print("peaks", omc_b200.measure_fp64_peak(), flush=True)
rng = np.random.default_rng(0)
for N, B in [(20, 8), (51, 8), (64, 8), (100, 148)]:
    V = rng.standard_normal((B, N, N)); V = V + np.transpose(V, (0, 2, 1))
    P, lam, sw, ms = omc_b200.psd_project_batch(V)
    err = 0.0; lerr = 0.0
    for b in range(B):
        l, Q = np.linalg.eigh(V[b]); Pe = (Q * np.maximum(l, 0)) @ Q.T
        err = max(err, np.abs(P[b] - Pe).max()); lerr = max(lerr, np.abs(np.sort(lam[b]) - l).max())
    print(f"psd N={N} B={B} P_err={err:.2e} lam_err={lerr:.2e} sweeps={sw[:4]} ms={ms:.3f}", flush=True)

for cfg in ["C1", "C2"]:
    k, A, mask, g = config_instance(cfg, 0)
    ct = CONFIGS[cfg]["cut_type"]
    p = Problem(k, A, mask, g, ct, state_pool_capacity=16)
    rp, ci, cp, ri = p.csr()
    orp, oci = M.mask_to_csr(mask); ocp, ori = M.mask_to_csc(mask)
    print(cfg, "csr exact", np.array_equal(rp, orp), np.array_equal(ci, oci), np.array_equal(cp, ocp), np.array_equal(ri, ori))
    X = rng.standard_normal(A.shape)
    got = p.objective_mse(X)
    want = (O.evaluate_objective(X, A, mask, X[:, :k], g), O.compute_MSE(X, A, mask, "in"), O.compute_MSE(X, A, mask, "out"), O.compute_MSE(X, A, mask, "all"))
    print(cfg, "objective/mse relerr", [abs(a - b) / abs(b) for a, b in zip(got, want)])
    opts = default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)
    oo = R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)
    t = time.time(); r = p.relax_batch([[]], opts)[0]; tg = time.time() - t
    t = time.time(); ro = R.solve_relaxation(A, mask, g, k, opts=oo); tc = time.time() - t
    print(cfg, "root gpu", r["status_code"], r["iters"], repr(r["objective"]), r["lower_bound"], r["res_p"], r["res_d"], f"{tg:.3f}s")
    print(cfg, "root cpu", ro["status"], ro["iters"], repr(ro["objective"]), ro["dual_objective"], f"{tc:.3f}s",
          "rel diff", abs(r["objective"] - ro["objective"]) / abs(ro["objective"]))
    print(cfg, "Y diff", np.abs(r["Y"] - ro["Y"]).max(), "X diff", np.abs(r["X"] - ro["X"]).max(), "U", np.abs(r["U"]).max())
    lam, vec, bp, feas = omc_b200.smallest_eigvecs_batch(r["Y"], r["U"], 1)
    xo, lo = E.breakpoint_vector(ro["Y"], ro["U"])
    print(cfg, "eig gpu", lam[0], "cpu", lo, "|cos|", abs(bp[0] @ xo), "feas", feas)
    # children with cuts (cuts built from the GPU root so both sides see identical inputs)
    x = bp[0]; Uh = r["U"]
    cid = p.add_cut(x, Uh)
    kids = [[Cut(cid, x, Uh, dirs)] for _, dirs in child_directions(ct, k)]
    t = time.time(); rk = p.relax_batch(kids, opts); tg = time.time() - t
    for (ind, dirs), rr in zip(child_directions(ct, k), rk):
        ro2 = R.solve_relaxation(A, mask, g, k, ct, [(x, Uh, dirs)], opts=oo)
        print(cfg, "child", dirs, "gpu", rr["status_code"], rr["iters"], repr(rr["objective"]), "cpu", ro2["status"], ro2["iters"], repr(ro2["objective"]),
              "rel", abs(rr["objective"] - ro2["objective"]) / abs(ro2["objective"]), flush=True)
    # throughput probe: 148*2 copies of the children batch
    nodes = (kids * 400)[: 296 if cfg == "C2" else 1184]
    f = p.frontier(nodes)
    ms = f.relax(opts); out = f.fetch(matrices=False); f.close()
    its = sum(o["iters"] for o in out)
    print(cfg, f"batch {len(nodes)} nodes kernel {ms:.2f} ms -> {len(nodes)/ms*1e3:.1f} nodes/s, total iters {its}, {ms*1e3/its*min(len(nodes),148):.2f} us/iter/CTA", flush=True)
    p.close()
print("DONE")
