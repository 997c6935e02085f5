"""rp components for nodes with cuts; Jacobi tolerance sweep; eigensolver re-check after the rewrite."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from omc_b200 import Problem, Cut, default_opts
from oracle.datagen import config_instance, CONFIGS
from oracle import relaxation as R
from oracle.cuts import child_directions

omc_b200.init(0)
rng = np.random.default_rng(0)
for N, B in [(20, 8), (51, 8), (100, 148)]:
    V = rng.standard_normal((B, N, N)); V = V + np.transpose(V, (0, 2, 1))
    P, lam, sw, ms = omc_b200.psd_project_batch(V)
    err = max(np.abs(P[b] - (lambda l, Q: (Q * np.maximum(l, 0)) @ Q.T)(*np.linalg.eigh(V[b]))).max() for b in range(B))
    print(f"psd N={N} B={B} P_err={err:.2e} sweeps={sw[:4]} ms={ms:.3f}", flush=True)
PH = ["wupd", "buildV", "gemm", "jacobi", "recon", "resid"]
for cfg in ["C1", "C2"]:
    k, A, mask, g = config_instance(cfg, 0)
    ct = CONFIGS[cfg]["cut_type"]
    p = Problem(k, A, mask, g, ct)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(A.shape[0]); x /= np.linalg.norm(x)
    Uh = 0.3 * rng.standard_normal((A.shape[0], k))
    cid = p.add_cut(x, Uh)
    dirs = child_directions(ct, k)[1][1]
    for mi in [1, 3, 100]:
        opts = default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adapt_every=0)
        f = p.frontier([[Cut(cid, x, Uh, dirs)]]); f.relax(opts); out = f.fetch(False); prof = f.profile(); f.close()
        print(cfg, dirs, "it", mi, "rp", out[0]["res_p"], "components psd1,psd2,psd3,trace,box,v,g:", prof[0, 8:15], flush=True)
    nodes = [[]] * 74 + [[Cut(cid, x, Uh, dirs)]] * 74
    for jt in [1e-7, 1e-5, 1e-3, 1e-2]:
        opts = default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=3000, jacobi_tol=jt)
        f = p.frontier(nodes); ms = f.relax(opts); out = f.fetch(False); prof = f.profile(); f.close()
        for nm, sl in [("root", slice(0, 74)), ("cut", slice(74, 148))]:
            pm = prof[sl].mean(axis=0); tot = pm[:6].sum()
            o = out[sl][0]
            print(cfg, nm, f"jtol {jt:g} kernel {ms:.1f} ms status {o['status_code']} iters {o['iters']} obj {o['objective']!r} sweeps/iter {pm[6]/pm[7]:.2f} cycles/iter {tot/pm[7]:.0f}",
                  " ".join(f"{PH[q]}={pm[q]/tot*100:.1f}%" for q in range(6)), flush=True)
    p.close()
print("DONE")
