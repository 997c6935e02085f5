"""The committed config-2 frontier fixture (tests/golden/c2_frontier.json): its node descriptors are valid cuts, and
the bounds the GPU engine recorded for them are reproduced by the CPU oracle (a CPU-side parity check of GPU output)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "c2_frontier.json")))


def test_fixture_descriptors_are_well_formed():
    fx = _load()
    assert fx["k"] == 1 and fx["n"] == 50 and fx["m"] == 50 and len(fx["nodes"]) == 64
    for nd in fx["nodes"]:
        assert len(nd["cuts"]) == nd["depth"] >= 1
        for c in nd["cuts"]:
            x = np.array(c["x"])
            assert x.shape == (50,) and abs(np.linalg.norm(x) - 1) < 1e-9 and abs(c["vhat"][0]) <= 1 + 1e-9
            assert c["dirs"][0] in ("left", "right")
        assert nd["gpu"]["objective"] >= nd["parent_bound"] * (1 - 1e-5)       # child bound >= parent bound


def test_oracle_reproduces_recorded_gpu_bounds():
    from oracle import relaxation as R
    from oracle.datagen import generate_matrix_completion_data
    fx = _load()
    A, mask = generate_matrix_completion_data(1, 50, 50, 1250, fx["seed"])
    done = 0
    for nd in sorted(fx["nodes"], key=lambda d: d["gpu"]["iters"]):
        if nd["gpu"]["status"] != 0 or nd["gpu"]["iters"] > 1500:
            continue
        cuts = [(np.array(c["x"]), np.array(c["vhat"]), c["dirs"]) for c in nd["cuts"]]
        r = R.solve_relaxation(A, mask, 80.0, 1, "linear", cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=5000))
        assert r["status"] == 0
        assert abs(r["objective"] - nd["gpu"]["objective"]) <= 1e-6 * abs(r["objective"]), (nd["node_id"], r["objective"], nd["gpu"]["objective"])
        done += 1
        if done == 2:
            break
    assert done >= 1
