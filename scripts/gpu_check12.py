"""C4 shape (k=3, 100x100, linear3): root + a few children, tracked vs exact, timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from oracle.datagen import config_instance
omc.init(0)
k, A, mask, g = config_instance("C4", 0)
p = omc.Problem(k, A, mask, g, "linear3")
names = ["wupd", "buildV", "gemm/lr", "jacobi", "recon", "resid"]
def run(nodes, label, **kw):
    f = omc.Frontier(p, nodes); ms = f.relax(omc.default_opts(**kw)); out = f.fetch(True); prof = f.profile(); f.close()
    pm = prof.sum(axis=0); tot = pm[:6].sum()
    print(f"{label}: {ms:.1f} ms nodes {len(nodes)} iters {pm[7]:.0f} cyc/iter {tot/pm[7]:.0f}", " ".join(f"{nm}={pm[q]/pm[7]/1e3:.1f}k" for q, nm in enumerate(names)),
          f"| lr {pm[14]:.0f} idle {pm[13]:.0f} full {pm[15]:.0f}", flush=True)
    return out
root_t = run([[]], "root tracked", max_iter=3000)[0]
root_e = run([[]], "root exact  ", max_iter=3000, exact_projection=1)[0]
print("root obj tracked %.9f exact %.9f it %d / %d st %d / %d" % (root_t["objective"], root_e["objective"], root_t["iters"], root_e["iters"], root_t["status_code"], root_e["status_code"]))
lam, vec, bp, feas = omc.smallest_eigvecs_batch(root_t["Y"][None], root_t["U"][None], 2)
from omc_b200.host import BBNode, create_matrix_cut_child_nodes
kids = create_matrix_cut_child_nodes(p, BBNode(node_id=1, parent_id=0, LB=root_t["objective"], depth=0), bp[0], root_t["U"], 1, root_t["objective"])
print("children", len(kids))
nodes = [kd.disjunctive_cuts for kd in kids]
ct = run(nodes, "64 children tracked", max_iter=3000)
ce = run(nodes[:8], "8 children exact", max_iter=3000, exact_projection=1)
for i in range(8):
    print(f"child {i}: tracked {ct[i]['objective']:.8f} it {ct[i]['iters']} st {ct[i]['status_code']} | exact {ce[i]['objective']:.8f} it {ce[i]['iters']} st {ce[i]['status_code']}")
