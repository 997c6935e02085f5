"""Same-box A/B: side-by-side tracking of the two large blocks on / off (OMC_NO_PAIR), 148 config-2 frontier nodes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200, bench
omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
cuts = bench.load_frontier_fixture(64)
nodes = [[omc_b200.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
big = (nodes * 3)[:148]
def run(tag):
    f = omc_b200.Frontier(p, big); ms = f.relax(omc_b200.default_opts(max_iter=5000)); out = f.fetch(False); prof = f.profile(); f.close()
    pm = prof.sum(axis=0)
    print(f"{tag}: {ms:.1f} ms iters {pm[7]:.0f} pair steps {pm[12]:.0f} obj0 {out[0]['objective']:.9f}", flush=True)
    return ms
res = {"pair": [], "nopair": []}
for rep in range(3):
    if not os.environ.get("ONLY_NOPAIR"):
        os.environ.pop("OMC_NO_PAIR", None); res["pair"].append(run("pair  "))
    os.environ["OMC_NO_PAIR"] = "1"; res["nopair"].append(run("nopair"))
print({k: (min(v), float(np.median(v))) for k, v in res.items() if v})
