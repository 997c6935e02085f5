"""Small deterministic run for ncu: 148 config-2 frontier nodes from the committed fixture, ONE launch of the fused
relaxation kernel with max_iter = 1500 (the launch to capture: `ncu -k regex:omc_relax_kernel -c 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200, bench
omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
cuts = bench.load_frontier_fixture(64)
nodes = [[omc_b200.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
nodes = (nodes * 3)[:148]
f = omc_b200.Frontier(p, nodes)
ms = f.relax(omc_b200.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=1500))
out = f.fetch(False); prof = f.profile()
pm = prof.sum(axis=0); tot = pm[:6].sum()
print(f"target launch: {ms:.2f} ms, {len(nodes)} nodes, iters {sum(o['iters'] for o in out)}, cycles/iter {tot/pm[7]:.0f}",
      " ".join(f"{nm}={pm[q]/tot*100:.1f}%" for q, nm in enumerate(["wupd", "buildV", "lowrank/gemm", "jacobi(full)", "recon", "resid"])))
