import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200, bench
omc_b200.init(0)
A, mask = bench.c2_instance(0)
p = omc_b200.Problem(1, A, mask, 80.0, "linear")
cuts = bench.load_frontier_fixture(64)
nodes = [[omc_b200.Cut(p.add_cut(x, vh), x, vh, d) for x, vh, d in cl] for cl in cuts]
big = (nodes * 19)[:1184]
for mi in (1, 1, 50):
    t0 = time.perf_counter(); f = omc_b200.Frontier(p, big); t1 = time.perf_counter()
    ms = f.relax(omc_b200.default_opts(max_iter=mi)); t2 = time.perf_counter()
    out = f.fetch(True); t3 = time.perf_counter(); f.close(); t4 = time.perf_counter()
    print(f"max_iter {mi}: create {1e3*(t1-t0):.1f} ms, relax {1e3*(t2-t1):.1f} ms (kernel {ms:.1f}), fetch {1e3*(t3-t2):.1f} ms, close {1e3*(t4-t3):.1f} ms", flush=True)
t0 = time.perf_counter(); r = p.relax_batch(big, omc_b200.default_opts(max_iter=1)); print(f"relax_batch total {1e3*(time.perf_counter()-t0):.1f} ms")
o = omc_b200.default_opts(max_iter=5000)
f = omc_b200.Frontier(p, big); ms1 = f.relax(o); ms2 = f.relax(o); out = f.fetch(False); f.close()
print(f"frontier.relax twice: {ms1:.0f} ms, {ms2:.0f} ms; mean iters {np.mean([r['iters'] for r in out]):.0f}", flush=True)
t0 = time.perf_counter(); r = p.relax_batch(big, o); t1 = time.perf_counter()
print(f"relax_batch: {1e3*(t1-t0):.0f} ms; mean iters {np.mean([q['iters'] for q in r]):.0f}", flush=True)
f = omc_b200.Frontier(p, big); ms3 = f.relax(o); f.close()
print(f"frontier.relax again: {ms3:.0f} ms", flush=True)
