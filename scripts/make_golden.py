"""Generates tests/golden/relax_golden.json with the CPU oracle (the reference ships no golden vectors).

Each case: a seeded instance, a list of cuts (seeded unit vectors x, fitted Uhat, directions) and the
oracle's relaxation bound at eps = 1e-9, together with its KKT certificate residuals.  The `-m gpu`
tests compare the CUDA path against these numbers without needing the oracle's solve at run time.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import relaxation as R
from oracle.datagen import generate_matrix_completion_data
from oracle.cuts import LABELS

CASES = [
    # name, k, n, m, n_indices, seed, cut_type, path = child index (0-based, reference enumeration order) taken at each level
    ("k1_10x10_root", 1, 10, 10, 50, 0, "linear", []),
    ("k1_10x10_c0", 1, 10, 10, 50, 0, "linear", [0]),
    ("k1_10x10_c1", 1, 10, 10, 50, 0, "linear", [1]),
    ("k1_10x10_c0c1", 1, 10, 10, 50, 0, "linear", [0, 1]),
    ("k1_10x10_c0c0c1", 1, 10, 10, 50, 0, "linear", [0, 0, 1]),
    ("k1_10x10_seed1_c1c0", 1, 10, 10, 50, 1, "linear", [1, 0]),
    ("k2_8x12_linear2_c0", 2, 8, 12, 60, 0, "linear2", [0]),
    ("k2_8x12_linear2_c2c5", 2, 8, 12, 60, 0, "linear2", [2, 5]),
    ("k2_8x12_linear3_c3", 2, 8, 12, 60, 0, "linear3", [3]),
    ("k2_8x12_linear3_c12c6", 2, 8, 12, 60, 0, "linear3", [12, 6]),
    ("k3_9x11_linear_c5", 3, 9, 11, 70, 2, "linear", [5]),
    ("k1_20x30_c1c1", 1, 20, 30, 300, 3, "linear", [1, 1]),
]
OPTS = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=200000)


def build_case(name, k, n, m, nidx, seed, cut_type, path):
    """Walks the B&B tree with the oracle: at each level the cut is (x, U) of the parent's relaxation
    (OMC.jl:2466-2468, 2522) and the child is path[level] in the reference's enumeration order."""
    from oracle.eigsep import breakpoint_vector
    from oracle.cuts import child_directions
    A, mask = generate_matrix_completion_data(k, n, m, nidx, seed)
    cuts = []
    for level, ci in enumerate(path):
        r = R.solve_relaxation(A, mask, 80.0, k, cut_type, cuts, opts=R.Options(**OPTS))
        x, _ = breakpoint_vector(r["Y"], r["U"], "smallest_1_eigvec")
        dirs = child_directions(cut_type, k)[ci][1]
        cuts.append((x, r["U"].copy(), dirs))
    return A, mask, cuts


def main():
    out = []
    for case in CASES:
        name, k, n, m, nidx, seed, cut_type, path = case
        A, mask, cuts = build_case(*case)
        r = R.solve_relaxation(A, mask, 80.0, k, cut_type, cuts, opts=R.Options(**OPTS))
        cert = R.certificate(r, A, mask, 80.0, k)
        out.append(dict(name=name, k=k, n=n, m=m, n_indices=nidx, seed=seed, gamma=80.0, cut_type=cut_type,
                        cuts=[dict(x=x.tolist(), Uhat=U.tolist(), dirs=d) for x, U, d in cuts],
                        status=int(r["status"]), iters=int(r["iters"]), objective=float(r["objective"]),
                        dual_objective=float(r["dual_objective"]), trY=float(np.trace(r["Y"])),
                        certificate={a: float(b) for a, b in cert.items()}))
        print(name, r["status"], r["iters"], r["objective"], r["dual_objective"], {a: f"{float(b):.1e}" for a, b in cert.items()})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "relax_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
