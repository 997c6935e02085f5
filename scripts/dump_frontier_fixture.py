"""Dumps tests/golden/c2_frontier.json: the cut descriptors (x, vhat = Uhat'x, directions) of the first 64 nodes of
bench.py's config-2 frontier (best-first expansion, incumbent withheld, built by the GPU engine) together with the
bound / status / iterations the GPU engine returned for them.  The descriptors are INPUT data for
bench.py --impl reference and the cpu_baseline leg (the CPU oracle cannot afford to expand a frontier in minutes);
tests/test_frontier_fixture.py re-solves some of them with the CPU oracle and checks the recorded GPU bounds.
Run on a GPU box:  python scripts/dump_frontier_fixture.py  (writes into gpurun_out/ and tests/golden/)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200
import bench

omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
nodes = bench.build_frontier_gpu(p, 1184, omc_b200)[:64]
res = p.relax_batch([nd.disjunctive_cuts for nd in nodes], omc_b200.default_opts(eps_abs=bench.EPS, eps_rel=bench.EPS, max_iter=bench.MAX_ITER))
out = dict(k=1, n=50, m=50, n_indices=1250, seed=0, gamma=80.0, cut_type="linear", eps=bench.EPS, max_iter=bench.MAX_ITER,
           nodes=[dict(node_id=nd.node_id, depth=nd.depth, parent_bound=nd.LB,
                       cuts=[dict(x=c.x.tolist(), vhat=(c.Uhat.T @ c.x).tolist(), dirs=c.directions) for c in nd.disjunctive_cuts],
                       gpu=dict(status=r["status_code"], iters=r["iters"], objective=r["objective"], lower_bound=r["lower_bound"]))
                  for nd, r in zip(nodes, res)])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (os.path.join(root, "gpurun_out", "c2_frontier.json"), os.path.join(root, "tests", "golden", "c2_frontier.json")):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"))
print("wrote", len(out["nodes"]), "nodes; iters", [n["gpu"]["iters"] for n in out["nodes"]][:16])
