import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, json
import omc_b200, bench
omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
cuts = bench.load_frontier_fixture(64)
nodes = [[omc_b200.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
names = ["wupd", "buildV", "gemm/lr", "jacobi", "recon", "resid"]
lrn = ["VZ", "resid", "cholqr", "VR", "gram", "jacobi", "combine", "(jacobi params)"]
def run(nodes_, label, **kw):
    f = omc_b200.Frontier(p, nodes_); ms = f.relax(omc_b200.default_opts(**kw)); out = f.fetch(False); prof = f.profile(); f.close()
    pm = prof.sum(axis=0); tot = pm[:6].sum()
    print(f"{label}: {ms:.1f} ms nodes {len(nodes_)} iters {pm[7]:.0f} cyc/iter {tot/pm[7]:.0f}", " ".join(f"{nm}={pm[q]/pm[7]/1e3:.1f}k" for q, nm in enumerate(names)),
          f"| lr proj {pm[14]:.0f} idle {pm[13]:.0f} full proj {pm[15]:.0f} sweeps {pm[6]:.0f}", flush=True)
    if pm[14] > 0:
        print("    lr step cycles/proj:", " ".join(f"{nm}={pm[16+q]/pm[14]/1e3:.2f}k" for q, nm in enumerate(lrn)), f"total={pm[16:23].sum()/pm[14]/1e3:.1f}k", flush=True)
        wn = ["X,T loops", "Y,U loops", "dense rows", "woodbury", "corrections"]
        print("    w-update cycles/iter:", " ".join(f"{nm}={pm[24+q]/pm[7]/1e3:.1f}k" for q, nm in enumerate(wn)), flush=True)
    return out
golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "check9_ref.json"))) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "check9_ref.json")) else None
sub = nodes[:4]
ol = run(sub, "lowrank ", max_iter=5000)
ref = [23.0132119088, 23.6438857896, 22.8287776975, 23.0952709181]
for i in range(len(sub)):
    print(f"node {i}: lowrank obj {ol[i]['objective']:.10f} it {ol[i]['iters']} st {ol[i]['status_code']}  relerr vs oracle {abs(ol[i]['objective']-ref[i])/abs(ref[i]):.2e}", flush=True)
big = (nodes * 3)[:148]
for mi in [200, 5000]:
    run(big, f"lowrank max_iter {mi}", max_iter=mi)
