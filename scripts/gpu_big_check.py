"""GPU diagnostic (round 2): the batched large-block engine (engine="batched") against its NumPy restatement
(oracle/bigblock.py), iterate for iterate and at convergence, and against the exact-projection oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import omc_b200 as omc
from oracle import relaxation as R, bigblock as Bg
from oracle.datagen import config_instance

omc.init(0)
which = sys.argv[1:] or ["traj", "conv"]

def traj(cfg, mis, cuts_fn=None):
    k, A, mask, g = config_instance(cfg, 0)
    ct = {"C1": "linear", "C2": "linear", "C3": "linear2", "C4": "linear3"}[cfg]
    p = omc.Problem(k, A, mask, g, ct)
    cuts = cuts_fn(A.shape[0], k) if cuts_fn else []
    gc = [omc.Cut(p.add_cut(x, Uh), x, Uh, d) for x, Uh, d in cuts]
    for mi in mis:
        o = omc.default_opts(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adapt_every=0)
        r = p.relax_batch([gc], o, engine="batched")[0]
        ro = Bg.solve_relaxation_big(A, mask, g, k, ct, cuts, opts=Bg.BigOptions(eps_abs=1e-30, eps_rel=1e-30, max_iter=mi, adaptive_rho=False))
        print(f"traj {cfg} L={len(cuts)} it={mi:4d}: dX {np.abs(r['X']-ro['X']).max():.2e} dY {np.abs(r['Y']-ro['Y']).max():.2e} dU {np.abs(r['U']-ro['U']).max():.2e}"
              f" | rp {r['res_p']:.6e} vs {ro['res_p']:.6e}  rd {r['res_d']:.6e} vs {ro['res_d']:.6e} | obj {r['objective']:.10f} vs {ro['objective']:.10f} lb {r['lower_bound']:.8f} vs {ro['lower_bound']:.8f} st {r['status_code']}", flush=True)
    p.close()

def chain(n, k, L=3, ct="linear3", seed=7):
    from oracle.cuts import LABELS
    rng = np.random.default_rng(seed)
    cuts = []
    for l in range(L):
        x = rng.standard_normal(n); x /= np.linalg.norm(x)
        Uh = 0.3 * rng.standard_normal((n, k))
        dirs = [LABELS[ct][rng.integers(len(LABELS[ct]) - 1)] for _ in range(k)]
        cuts.append((x, Uh, dirs))
    return cuts

def conv(cfg, cuts_fn=None, exact=True, eps=1e-8):
    k, A, mask, g = config_instance(cfg, 0)
    ct = {"C1": "linear", "C2": "linear", "C3": "linear2", "C4": "linear3"}[cfg]
    p = omc.Problem(k, A, mask, g, ct)
    cuts = cuts_fn(A.shape[0], k) if cuts_fn else []
    gc = [omc.Cut(p.add_cut(x, Uh), x, Uh, d) for x, Uh, d in cuts]
    o = omc.default_opts(eps_abs=eps, eps_rel=eps, max_iter=8000)
    t0 = time.time()
    f = p.frontier([gc], engine="batched"); ms = f.relax(o); r = f.fetch()[0]; stt = f.stats(); f.close()
    t1 = time.time()
    ro = Bg.solve_relaxation_big(A, mask, g, k, ct, cuts, opts=Bg.BigOptions(eps_abs=eps, eps_rel=eps, max_iter=8000))
    msg = f"conv {cfg} L={len(cuts)}: gpu it {r['iters']} st {r['status_code']} obj {r['objective']:.10f} lb {r['lower_bound']:.8f} ({ms:.0f} ms, {stt['launches']} launches) | numpy-big it {ro['iters']} st {ro['status']} obj {ro['objective']:.10f} lb {ro['lower_bound']:.8f}"
    if exact:
        re_ = R.solve_relaxation(A, mask, g, k, ct, cuts, opts=R.Options(eps_abs=eps, eps_rel=eps, max_iter=8000))
        msg += f" | exact it {re_['iters']} st {re_['status']} obj {re_['objective']:.10f} rel {abs(r['objective']-re_['objective'])/abs(re_['objective']):.1e}"
    print(msg, flush=True)
    p.close()

def guarded(fn, *a, **kw):
    try:
        fn(*a, **kw)
    except Exception as e:
        print("FAILED", fn.__name__, a[:1], repr(e)[:200], flush=True)

if "traj" in which:
    guarded(traj, "C2", [1, 2, 3])
    guarded(traj, "C3", [1, 2], lambda n, k: chain(n, k, 2, "linear2"))
    guarded(traj, "C4", [1, 2])
    guarded(traj, "C4", [1, 2], lambda n, k: chain(n, k, 3, "linear3"))
if "conv" in which:
    guarded(conv, "C2"); guarded(conv, "C3"); guarded(conv, "C3", lambda n, k: chain(n, k, 2, "linear2"))
    guarded(conv, "C4", exact=True, eps=1e-9); guarded(conv, "C4", lambda n, k: chain(n, k, 3, "linear3"), exact=True, eps=1e-9)
    guarded(conv, "C4", lambda n, k: chain(n, k, 8, "linear3", seed=11), exact=True, eps=1e-9)
if "batch" in which:
    # a small batch of identical and different nodes: batch == singles
    k, A, mask, g = config_instance("C4", 0)
    p = omc.Problem(k, A, mask, g, "linear3")
    sets = [[], chain(100, 3, 1), chain(100, 3, 2, seed=3), chain(100, 3, 4, seed=5), []]
    gcs = [[omc.Cut(p.add_cut(x, Uh), x, Uh, d) for x, Uh, d in cs] for cs in sets]
    o = omc.default_opts(eps_abs=1e-8, eps_rel=1e-8, max_iter=4000)
    f = p.frontier(gcs, engine="batched"); ms = f.relax(o); rb = f.fetch(); stt = f.stats(); f.close()
    print("batch of 5:", [(r["iters"], r["status_code"], round(r["objective"], 8)) for r in rb], f"{ms:.0f} ms", stt)
    for i, gc in enumerate(gcs):
        r1 = p.relax_batch([gc], o, engine="batched")[0]
        print("  single", i, r1["iters"], r1["status_code"], round(r1["objective"], 8), "d", abs(r1["objective"] - rb[i]["objective"]))
    p.close()
