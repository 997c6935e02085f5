"""Builds tests/golden/<cfg>_frontier_pool.json: ONE frontier of open branch-and-bound nodes of a BASELINE configuration
(best-first disjunctive expansion by the GPU engine with the root heuristic's incumbent as cut-off, exactly what the
branch-and-bound loop does), stored in pool form (every cut once, nodes as (pool id, direction codes)), plus the GPU engine's
result for the first 64 nodes.  INPUT data for bench.py (both arms relax nodes of this one frontier; at N GPUs it is sharded, not
rebuilt) and for tests/test_frontier_fixture.py.  Run on a GPU box:  OMC_BENCH_CFG=C4 python scripts/dump_frontier_pool.py 9472"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
import bench
from omc_b200.engine import LABELS

target = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
W = bench.W
omc.init(0)
A, mask = bench.instance(0)
p = omc.Problem(W["k"], A, mask, bench.GAMMA, W["ct"])
# incumbent: the reference's root heuristic (OMC.jl:521-621) -- zero-filled SVD start, alternating minimisation, objective
U0 = np.linalg.svd(np.where(mask, A, 0.0))[0][:, :W["k"]]
am = omc.alternating_minimization(p, U0)
incumbent = p.objective_mse(am["U"] @ am["V"])[0]
print("incumbent", incumbent, "alt-min sweeps", am["n_iters"], flush=True)
t0 = time.time()
nodes = bench.build_frontier_gpu(p, target, omc, cutoff=incumbent)
print("frontier of", len(nodes), "nodes in %.0f s" % (time.time() - t0), "depths", np.bincount([nd.depth for nd in nodes]).tolist(), flush=True)
lab = LABELS[W["ct"]]
pool_index, pool = {}, []
out_nodes = []
for nd in nodes:
    ent = []
    for c in nd.disjunctive_cuts:
        if c.cut_id not in pool_index:
            pool_index[c.cut_id] = len(pool)
            pool.append([np.asarray(c.x).tolist(), (np.asarray(c.Uhat).T @ c.x).tolist()])
        ent.append([pool_index[c.cut_id]] + [lab.index(d) for d in c.directions])
    out_nodes.append(ent)
first = nodes[:64]
res = p.relax_batch([nd.disjunctive_cuts for nd in first], omc.default_opts(eps_abs=bench.EPS, eps_rel=bench.EPS, max_iter=bench.MAX_ITER, cutoff=incumbent))
out = dict(cfg=bench.CFG, k=W["k"], n=W["n"], m=W["m"], n_indices=W["nidx"], seed=0, gamma=bench.GAMMA, cut_type=W["ct"], eps=bench.EPS,
           max_iter=bench.MAX_ITER, incumbent=incumbent, pool=pool, nodes=out_nodes,
           gpu_first=[dict(node_id=nd.node_id, depth=nd.depth, parent_bound=nd.LB, status=r["status_code"], iters=r["iters"], objective=r["objective"],
                           lower_bound=r["lower_bound"]) for nd, r in zip(first, res)])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = f"{bench.CFG.lower()}_frontier_pool.json"
for path in (os.path.join(root, "gpurun_out", name), os.path.join(root, "tests", "golden", name)):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"))
print("wrote", len(out_nodes), "nodes,", len(pool), "pool cuts; first statuses", [r["status_code"] for r in res][:32], "iters", [r["iters"] for r in res][:16])
