import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, json
import omc_b200, bench
omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
cuts = bench.load_frontier_fixture(64)
nodes = [[omc_b200.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
nodes = (nodes * 3)[:148]
for mi in [200, 1000, 5000]:
    f = omc_b200.Frontier(p, nodes); ms = f.relax(omc_b200.default_opts(max_iter=mi)); out = f.fetch(False); prof = f.profile(); f.close()
    pm = prof.sum(axis=0); tot = pm[:6].sum()
    print(f"max_iter {mi}: {ms:.1f} ms iters {pm[7]:.0f} sweeps/iter {pm[6]/pm[7]:.2f} cyc/iter {tot/pm[7]:.0f}", " ".join(f"{nm}={pm[q]/tot*100:.1f}%" for q, nm in enumerate(["wupd", "buildV", "gemm", "jacobi", "recon", "resid"])),
          f"| blk1: rot/call {pm[8]/pm[10]:.0f} blkcyc/call {pm[9]/pm[10]:.0f} | blk2+3: rot/call {pm[11]/pm[13]:.0f} blkcyc/call {pm[12]/pm[13]:.0f}", flush=True)
