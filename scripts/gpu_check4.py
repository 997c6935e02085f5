"""Frontier workload statistics for the bench (C2): iteration histogram, statuses, throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
from omc_b200 import Problem, default_opts, expand_frontier
from oracle.datagen import config_instance, CONFIGS
omc_b200.init(0)
PH = ["wupd", "buildV", "gemm", "jacobi", "recon", "resid"]
for cfg, target in [("C1", 592), ("C2", 592)]:
    k, A, mask, g = config_instance(cfg, 0)
    p = Problem(k, A, mask, g, CONFIGS[cfg]["cut_type"])
    u, s, vt = np.linalg.svd(np.where(mask, A, 0.0)); Xr = (u[:, :k] * s[:k]) @ vt[:k]
    ub = p.objective_mse(Xr)[0]
    root = p.relax_batch([[]])[0]
    print(cfg, "root", root["objective"], "svd ub", ub, flush=True)
    for cutoff in [float("inf"), ub]:
        t = time.time()
        nodes, st = expand_frontier(p, target, default_opts(max_iter=5000), cutoff=cutoff)
        print(cfg, "cutoff", cutoff, "frontier", len(nodes), "depths", sorted(set(n.depth for n in nodes)), st, f"{time.time()-t:.1f}s", flush=True)
        if not nodes: continue
        for co in [float("inf"), ub]:
            f = p.frontier([n.disjunctive_cuts for n in nodes])
            ms = f.relax(default_opts(max_iter=5000, cutoff=co)); out = f.fetch(False); prof = f.profile(); f.close()
            its = np.array([o["iters"] for o in out]); stc = np.bincount([o["status_code"] for o in out], minlength=5)
            pm = prof.sum(axis=0); tot = pm[:6].sum()
            print(cfg, f"  relax cutoff={co:.4g}: {ms:.1f} ms -> {len(nodes)/ms*1e3:.1f} nodes/s; iters mean {its.mean():.0f} med {np.median(its):.0f} max {its.max()} status {stc}",
                  f"sweeps/iter {pm[6]/pm[7]:.2f} cyc/iter {tot/pm[7]:.0f}", " ".join(f"{PH[q]}={pm[q]/tot*100:.0f}%" for q in range(6)), flush=True)
    p.close()
print("DONE")
