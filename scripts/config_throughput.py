"""Per-config relaxation throughput on one B200 (BASELINE.json configs 1-4 without the Shor rows): a breadth-first frontier
of open nodes is built untimed, then relaxed in ONE launch from cold starts.  Writes profiles/r01_configs_summary.json."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from omc_b200.host import expand_frontier
from oracle.datagen import config_instance, CONFIGS
omc.init(0)
out = {}
plan = {"C1": dict(target=592, max_iter=5000, build_iter=3000), "C2": dict(target=148, max_iter=5000, build_iter=3000),
        "C3": dict(target=148, max_iter=5000, build_iter=2000), "C4": dict(target=8, max_iter=600, build_iter=600)}
for cfg, pl in plan.items():
    k, A, mask, g = config_instance(cfg, 0)
    ct = CONFIGS[cfg]["cut_type"]
    nev = 2 if cfg == "C4" else 1
    p = omc.Problem(k, A, mask, g, ct)
    t0 = time.time()
    nodes, st = expand_frontier(p, pl["target"], omc.default_opts(max_iter=pl["build_iter"]), nev=nev, max_levels=12)
    t_build = time.time() - t0
    if not nodes:
        print(cfg, "no open nodes", st); p.close(); continue
    f = omc.Frontier(p, [nd.disjunctive_cuts for nd in nodes])
    ms = f.relax(omc.default_opts(max_iter=pl["max_iter"]))
    res = f.fetch(False); prof = f.profile(); f.close()
    pm = prof.sum(axis=0)
    status = np.bincount([r["status_code"] for r in res], minlength=5).tolist()
    n, m = A.shape
    flops_iter = 16.0 / 3.0 * ((n + m) ** 3 + (n + k) ** 3 + n ** 3)
    out[cfg] = {"shape": [int(k), int(n), int(m)], "cut_type": ct, "open_nodes": len(nodes), "depth_range": [min(nd.depth for nd in nodes), max(nd.depth for nd in nodes)],
                "kernel_ms": ms, "nodes_per_s": len(nodes) / (ms * 1e-3), "iters_mean": float(pm[7] / len(nodes)), "max_iter": pl["max_iter"],
                "status_counts[opt,iterlim,infeas,time,cutoff]": status,
                "projections": {"tracked": float(pm[14]), "idle": float(pm[13]), "full": float(pm[15])},
                "algorithmic_tflops": float(pm[7]) * flops_iter / (ms * 1e-3) * 1e-12, "frontier_build_s": t_build}
    print(cfg, json.dumps(out[cfg]), flush=True)
    p.close()
os.makedirs(os.path.join(os.path.dirname(__file__), "..", "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "configs_summary.json"), "w"), indent=1)
