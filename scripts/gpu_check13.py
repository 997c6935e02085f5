"""Warm start from the parent's record (state + tracked bases) vs cold start, C2 frontier fixture nodes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc, bench
from oracle.datagen import config_instance
omc.init(0)
k, A, mask, g = config_instance("C2", 0)
NB = 64
p = omc.Problem(k, A, mask, g, "linear", state_pool_capacity=NB)
cuts = bench.load_frontier_fixture(NB)
nodes = [[omc.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
parents = [nd[:-1] for nd in nodes]
names = ["wupd", "buildV", "gemm/lr", "jacobi", "recon", "resid"]
def run(nl, label, warm=None, save=None, **kw):
    f = omc.Frontier(p, nl, warm, save); ms = f.relax(omc.default_opts(**kw)); out = f.fetch(False); prof = f.profile(); f.close()
    pm = prof.sum(axis=0); tot = pm[:6].sum()
    print(f"{label}: {ms:.1f} ms nodes {len(nl)} iters {pm[7]:.0f} (mean {pm[7]/len(nl):.0f}) cyc/iter {tot/pm[7]:.0f}", " ".join(f"{nm}={pm[q]/pm[7]/1e3:.1f}k" for q, nm in enumerate(names)),
          f"| lr {pm[14]:.0f} idle {pm[13]:.0f} full {pm[15]:.0f}", flush=True)
    return out
pr = run(parents, "parents cold (saved)", save=list(range(NB)), max_iter=5000)
cw = run(nodes, "children warm       ", warm=list(range(NB)), max_iter=5000)
cc = run(nodes, "children cold       ", max_iter=5000)
d = [abs(a["objective"] - b["objective"]) / abs(b["objective"]) for a, b in zip(cw, cc) if a["status_code"] == 0 and b["status_code"] == 0]
print("warm vs cold objective rel diff: max %.2e median %.2e over %d converged pairs; status warm %s cold %s" % (max(d), np.median(d), len(d), np.bincount([a["status_code"] for a in cw], minlength=2), np.bincount([a["status_code"] for a in cc], minlength=2)))
