"""altmin parity vs oracle + rerun of GPU tests."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from omc_b200 import Problem, Cut
from oracle.datagen import config_instance, CONFIGS
from oracle import altmin as AM
from oracle.cuts import LABELS
omc_b200.init(0)
for cfg in ["C1", "C2", "C3", "C4"]:
    k, A, mask, g = config_instance(cfg, 0); n = A.shape[0]; ct = CONFIGS[cfg]["cut_type"]
    p = Problem(k, A, mask, g, ct)
    U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :k]
    for name, Ui in [("U0", U0), ("-U0", -U0)]:
        t = time.time(); r = omc_b200.alternating_minimization(p, Ui); tg = time.time() - t
        t = time.time(); ro = AM.alternating_minimization(A, n, k, mask, g, True, ct, Ui); tc = time.time() - t
        print(cfg, name, "gpu", r["converged"], r["n_iters"], r["objectives"][-1], f"{r['solve_time']*1e3:.1f}ms", "cpu", ro["converged"], ro["n_iters"], ro["objectives"][-1], f"{tc*1e3:.0f}ms",
              "dU", np.abs(r["U"] - ro["U"]).max(), "dV", np.abs(r["V"] - ro["V"]).max(), flush=True)
    rng = np.random.default_rng(0); x = rng.standard_normal(n); x /= np.linalg.norm(x); Uh = 0.3 * rng.standard_normal((n, k))
    dirs = [LABELS[ct][-1]] * k
    cid = p.add_cut(x, Uh)
    r = omc_b200.alternating_minimization(p, U0, [Cut(cid, x, Uh, dirs)])
    ro = AM.alternating_minimization(A, n, k, mask, g, True, ct, U0, [(x, Uh, dirs)])
    print(cfg, "cut", "gpu", r["converged"], r["n_iters"], r["objectives"][-1], "cpu", ro["converged"], ro["n_iters"], ro["objectives"][-1], "dU", np.abs(r["U"] - ro["U"]).max(), flush=True)
    p.close()
print("DONE")
