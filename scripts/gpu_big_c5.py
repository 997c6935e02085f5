"""GPU check (round 2): the batched engine at config-5 size (k = 5, 1000 x 1000, PSD blocks 2000 / 1005 / 1000).
(1) KAT-root-full: every entry observed -> closed-form bound (oracle/kat.py); (2) the config-5 root (20 % observed);
(3) its 32 children in one lockstep batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from omc_b200.synthetic import generate_matrix_completion_data
from omc_b200.host import BBNode, create_matrix_cut_child_nodes
from oracle import kat
omc.init(0)
which = sys.argv[1:] or ["kat", "root", "kids"]
k, n, m = 5, 1000, 1000
if "kat" in which:
    A, _ = generate_matrix_completion_data(k, n, m, n * m, 1)
    mask = np.ones((n, m), bool)
    want = kat.root_bound_full(A, 80.0, k)
    p = omc.Problem(k, A, mask, 80.0, "linear")
    f = p.frontier([[]], engine="batched"); ms = f.relax(omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)); r = f.fetch()[0]; st = f.stats(); f.close()
    print(f"KAT full 1000x1000 k=5: it {r['iters']} st {r['status_code']} obj {r['objective']:.8f} closed form {want:.8f} rel {abs(r['objective']-want)/want:.2e} lb {r['lower_bound']:.6f}; {ms:.0f} ms, {ms*1e3/st['iterations']:.0f} us/iteration, node bytes {st['node_bytes']/1e6:.1f} MB", flush=True)
    p.close()
if "root" in which or "kids" in which:
    A, mask = generate_matrix_completion_data(k, n, m, 200000, 0)
    p = omc.Problem(k, A, mask, 80.0, "linear")
    o = omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000)
    f = p.frontier([[]], engine="batched"); ms = f.relax(o); r = f.fetch()[0]; st = f.stats(); f.close()
    print(f"C5 root: it {r['iters']} st {r['status_code']} obj {r['objective']:.8f} lb {r['lower_bound']:.6f}; {ms:.0f} ms, {ms*1e3/st['iterations']:.0f} us/iteration", flush=True)
    if "kids" in which:
        lam, vec, bp, feas = omc.smallest_eigvecs_batch(r["Y"], r["U"], 1)
        kids = create_matrix_cut_child_nodes(p, BBNode(1, 0, -np.inf, 0), bp[0], r["U"], 1, r["objective"])
        f = p.frontier([nd.disjunctive_cuts for nd in kids], engine="batched"); ms = f.relax(o); res = f.fetch(matrices=False); st = f.stats(); f.close()
        its = np.array([x["iters"] for x in res]); sc = np.bincount([x["status_code"] for x in res], minlength=6)
        print(f"C5 children: B={len(kids)} {ms:.0f} ms; status {sc.tolist()} iters mean {its.mean():.0f} max {its.max()}; {ms*1e3/st['iterations']:.0f} us per lockstep iteration, {ms*1e3/st['node_iterations']:.1f} us per node-iteration; objs {np.round([x['objective'] for x in res][:6], 4).tolist()}", flush=True)
    p.close()
