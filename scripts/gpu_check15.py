"""Cold phase only (max_iter = 14): warm eigenbasis pre-rotation vs restart from the identity, Jacobi tolerance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200, bench
omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
cuts = bench.load_frontier_fixture(64)
nodes = [[omc_b200.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
big = (nodes * 3)[:148]
names = ["wupd", "buildV", "gemm/lr", "jacobi", "recon", "resid"]
def run(label, **kw):
    f = omc_b200.Frontier(p, big); ms = f.relax(omc_b200.default_opts(**kw)); out = f.fetch(False); prof = f.profile(); f.close()
    pm = prof.sum(axis=0); tot = pm[:6].sum()
    print(f"{label}: {ms:.1f} ms iters {pm[7]:.0f} cyc/iter {tot/pm[7]:.0f}", " ".join(f"{nm}={pm[q]/pm[7]/1e3:.1f}k" for q, nm in enumerate(names)),
          f"| lr {pm[14]:.0f} idle {pm[13]:.0f} full {pm[15]:.0f} sweeps {pm[6]:.0f} sweeps/full {pm[6]/max(1,pm[15]):.2f}", flush=True)
for mi in (14, 40):
    run(f"max_iter {mi} default       ", max_iter=mi)
    run(f"max_iter {mi} reortho_every=1", max_iter=mi, reortho_every=1)
    run(f"max_iter {mi} jacobi_tol 1e-3", max_iter=mi, jacobi_tol=1e-3)
    run(f"max_iter {mi} jacobi_tol 1e-7", max_iter=mi, jacobi_tol=1e-7)
