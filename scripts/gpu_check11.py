import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc, bench
from oracle.datagen import config_instance
omc.init(0)
k, A, mask, g = config_instance("C2", 0)
p = omc.Problem(k, A, mask, g, "linear")
cuts = bench.load_frontier_fixture(6)
nodes = [[omc.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
for eps in (1e-8, 1e-9):
    f = omc.Frontier(p, nodes); f.relax(omc.default_opts(max_iter=9000, eps_abs=eps, eps_rel=eps)); rt = f.fetch(False); prof = f.profile(); f.close()
    re_ = p.relax_batch(nodes, omc.default_opts(max_iter=9000, exact_projection=1, eps_abs=eps, eps_rel=eps))
    for i, (a, b) in enumerate(zip(rt, re_)):
        print(f"eps {eps:g} node {i}: tracked it {a['iters']} st {a['status_code']} obj {a['objective']:.9f} lb {a['lower_bound']:.9f} | exact it {b['iters']} st {b['status_code']} obj {b['objective']:.9f} lb {b['lower_bound']:.9f} | rel diff {abs(a['objective']-b['objective'])/abs(b['objective']):.2e}  lr {prof[i,14]:.0f} idle {prof[i,13]:.0f} full {prof[i,15]:.0f}", flush=True)
