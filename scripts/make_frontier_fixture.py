"""Builds tests/golden/c2_frontier.json with the CPU oracle: the cut descriptors of the first open nodes of the
best-first expansion (incumbent withheld) of BASELINE config 2, seed 0.  bench.py --impl reference and the
cpu_baseline leg time the oracle on these nodes; the GPU tests relax them and compare with the stored bounds."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np
from oracle.datagen import generate_matrix_completion_data

EPS, MAX_ITER = 1e-8, 5000


def work(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import relaxation as R
    from oracle.eigsep import breakpoint_vector, master_feasible
    A, mask, cuts = args
    t = time.time()
    r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts, opts=R.Options(eps_abs=EPS, eps_rel=EPS, max_iter=MAX_ITER))
    x, _ = breakpoint_vector(r["Y"], r["U"])
    return dict(status=int(r["status"]), iters=int(r["iters"]), objective=float(r["objective"]), feas=bool(master_feasible(r["Y"], r["U"])),
                x=x, U=r["U"].copy(), seconds=time.time() - t)


def main(target=32):
    A, mask = generate_matrix_completion_data(1, 50, 50, 1250, 0)
    open_nodes = [(-np.inf, 1, [])]
    counter, relaxed = 1, []
    with mp.get_context("fork").Pool(8) as pool:
        while len(open_nodes) < target:
            open_nodes.sort(key=lambda t: (t[0], t[1]))
            batch = [open_nodes.pop(0) for _ in range(min(8, len(open_nodes), target - len(open_nodes)))]
            res = pool.map(work, [(A, mask, c) for _, _, c in batch])
            for (lb, nid, cuts), r in zip(batch, res):
                print(nid, len(cuts), r["status"], r["iters"], r["objective"], r["feas"], f"{r['seconds']:.1f}s", flush=True)
                relaxed.append(dict(node_id=nid, depth=len(cuts), status=r["status"], iters=r["iters"], objective=r["objective"]))
                if r["status"] != 0 or r["feas"]:
                    continue
                for ind, d in ((1, ["left"]), (2, ["right"])):
                    open_nodes.append((r["objective"], counter + ind, cuts + [(r["x"], r["U"], d)]))
                counter += 2
    open_nodes.sort(key=lambda t: (t[0], t[1]))
    out = dict(k=1, n=50, m=50, n_indices=1250, seed=0, gamma=80.0, cut_type="linear", eps=EPS, max_iter=MAX_ITER, relaxed=relaxed,
               nodes=[dict(node_id=nid, parent_bound=lb, cuts=[dict(x=x.tolist(), vhat=(U.T @ x).tolist(), dirs=d) for x, U, d in cuts])
                      for lb, nid, cuts in open_nodes[:target]])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c2_frontier.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, len(out["nodes"]))


if __name__ == "__main__":
    main()
