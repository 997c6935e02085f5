"""Small deterministic run for ncu: builds a 148-node C2 frontier, then launches the fused relaxation kernel once with
max_iter = 200 (the launch to capture).  Prints relax_launches_before=N for `ncu -k regex:omc_relax_kernel -s N -c 1`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
import bench

omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
launches = [0]
orig = p.relax_batch
def counted(*a, **k):
    launches[0] += 1
    return orig(*a, **k)
p.relax_batch = counted
nodes = bench.build_frontier_gpu(p, 148, omc_b200)
print(f"relax_launches_before={launches[0]}", flush=True)
f = omc_b200.Frontier(p, [nd.disjunctive_cuts for nd in nodes])
ms = f.relax(omc_b200.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=200))
out = f.fetch(False); prof = f.profile()
pm = prof.sum(axis=0); tot = pm[:6].sum()
print(f"target launch: {ms:.2f} ms, {len(nodes)} nodes, iters {sum(o['iters'] for o in out)}, sweeps/iter {pm[6]/pm[7]:.2f}, cycles/iter {tot/pm[7]:.0f}",
      " ".join(f"{nm}={pm[q]/tot*100:.1f}%" for q, nm in enumerate(["wupd", "buildV", "gemm", "jacobi", "recon", "resid"])))
