"""Localise the L>0 divergence: GPU vs oracle after a fixed number of ADMM iterations; kernel phase profile."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from omc_b200 import Problem, Cut, default_opts
from oracle.datagen import config_instance, CONFIGS
from oracle import relaxation as R
from oracle.cuts import child_directions

omc_b200.init(0)
PH = ["wupd", "buildV", "gemm", "jacobi", "recon", "resid"]
for cfg in ["C1", "C2"]:
    k, A, mask, g = config_instance(cfg, 0)
    ct = CONFIGS[cfg]["cut_type"]
    p = Problem(k, A, mask, g, ct)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(A.shape[0]); x /= np.linalg.norm(x)
    Uh = 0.3 * rng.standard_normal((A.shape[0], k))
    cid = p.add_cut(x, Uh)
    for dirs in [d for _, d in child_directions(ct, k)]:
        for mi in [1, 2, 3, 5, 10, 25, 100]:
            opts = default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adapt_every=0)
            oo = R.Options(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adaptive_rho=False)
            r = p.relax_batch([[Cut(cid, x, Uh, dirs)]], opts)[0]
            ro = R.solve_relaxation(A, mask, g, k, ct, [(x, Uh, dirs)], opts=oo)
            print(cfg, dirs, "it", mi, "dX %.2e dY %.2e dU %.2e" % (np.abs(r["X"] - ro["X"]).max(), np.abs(r["Y"] - ro["Y"]).max(), np.abs(r["U"] - ro["U"]).max()),
                  "obj", r["objective"], ro["objective"], "rp %.2e/%.2e rd %.2e/%.2e" % (r["res_p"], ro["res_p"], r["res_d"], ro["res_d"]), flush=True)
    # profile: root nodes x grid, and the cut node
    for name, nodes in [("root", [[]] * 148), ("cut", [[Cut(cid, x, Uh, child_directions(ct, k)[0][1])]] * 148)]:
        opts = default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=400)
        f = p.frontier(nodes); ms = f.relax(opts); out = f.fetch(matrices=False); prof = f.profile(); f.close()
        pm = prof.mean(axis=0)
        tot = pm[:6].sum()
        print(cfg, name, f"kernel {ms:.2f} ms iters {pm[7]:.0f} sweeps/iter {pm[6]/pm[7]:.2f} cycles/iter {tot/pm[7]:.0f}",
              " ".join(f"{PH[q]}={pm[q]/tot*100:.1f}%" for q in range(6)), "status", out[0]["status_code"], flush=True)
    p.close()
print("DONE")
