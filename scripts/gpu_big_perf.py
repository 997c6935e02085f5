"""GPU measurement (round 2): throughput of the batched engine on a config-4 frontier (k=3, 100x100, linear3, smallest_2_eigvec)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from omc_b200.synthetic import generate_matrix_completion_data
from omc_b200.host import BBNode, create_matrix_cut_child_nodes

omc.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
k, n, m = 3, 100, 100
A, mask = generate_matrix_completion_data(k, n, m, 3000, 0)
p = omc.Problem(k, A, mask, 80.0, "linear3")
opts = omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=max_iter)
t0 = time.time()
root = BBNode(1, 0, -np.inf, 0)
r = p.relax_batch([[]], opts)[0]
lam, vec, bp, feas = omc.smallest_eigvecs_batch(r["Y"], r["U"], 2)
level = create_matrix_cut_child_nodes(p, root, bp[0], r["U"], 1, r["objective"])
counter = 1 + len(level)
print("root", r["iters"], r["termination_status"], r["objective"], "children", len(level), flush=True)
while len(level) < B:
    f = p.frontier([nd.disjunctive_cuts for nd in level]); ms = f.relax(opts); res = f.fetch(); st = f.stats(); f.close()
    sc = np.bincount([x["status_code"] for x in res], minlength=6)
    print(f"level of {len(level)}: {ms:.0f} ms, status {sc.tolist()}, iters mean {np.mean([x['iters'] for x in res]):.0f}, stats {st}", flush=True)
    ok = [i for i, x in enumerate(res) if x["status_code"] == 0]
    ok.sort(key=lambda i: res[i]["objective"])
    ok = ok[: max(1, (B + 62) // 63)]
    lam, vec, bp, feas = omc.smallest_eigvecs_batch(np.stack([res[i]["Y"] for i in ok]), np.stack([res[i]["U"] for i in ok]), 2)
    nxt = []
    for q, i in enumerate(ok):
        if feas[q]:
            continue
        kids = create_matrix_cut_child_nodes(p, level[i], bp[q], res[i]["U"], counter, res[i]["objective"])
        counter += len(kids); nxt += kids
    level = [nd for j, nd in enumerate(level) if j not in set(ok)] + nxt
level = level[:B]
print("frontier built", len(level), "nodes in %.1fs" % (time.time() - t0), "cuts/node", np.mean([len(nd.disjunctive_cuts) for nd in level]), flush=True)
f = p.frontier([nd.disjunctive_cuts for nd in level])
for rep in range(2):
    ms = f.relax(opts); res = f.fetch(matrices=False); st = f.stats()
    sc = np.bincount([x["status_code"] for x in res], minlength=6)
    its = np.array([x["iters"] for x in res])
    print(f"rep {rep}: B={B} {ms:.0f} ms -> {B / ms * 1e3:.1f} nodes/s; status[opt,iterlim,infeas,time,cutoff,num] {sc.tolist()}; iters mean {its.mean():.0f} max {its.max()}; "
          f"lockstep iterations {st['iterations']}, node-iterations {st['node_iterations']}, launches {st['launches']}, us per lockstep iteration {ms * 1e3 / st['iterations']:.0f}, "
          f"ns per node-iteration {ms * 1e6 / st['node_iterations']:.0f}", flush=True)
f.close(); p.close()
