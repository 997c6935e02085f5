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


def _load_pool(cfg):
    return json.load(open(os.path.join(ROOT, "tests", "golden", f"{cfg}_frontier_pool.json")))


def test_pool_fixtures_are_well_formed():
    """The committed config-4 / config-5 frontiers (pool form, scripts/dump_frontier_pool.py): every cut is a unit breakpoint
    vector with |vhat| <= 1, every node a list of (pool id, k direction codes), and the recorded GPU bounds lie between the
    parent's bound and the incumbent."""
    for cfg, nn, npool in (("c4", 9472, 157), ("c5", 1024, 33)):
        fx = _load_pool(cfg)
        k, n = fx["k"], fx["n"]
        ndir = {"linear": 2, "linear2": 3, "linear3": 4}[fx["cut_type"]]
        assert len(fx["nodes"]) == nn and len(fx["pool"]) == npool and np.isfinite(fx["incumbent"])
        for x, vh in fx["pool"]:
            assert len(x) == n and len(vh) == k and abs(np.linalg.norm(x) - 1) < 1e-9 and np.abs(vh).max() <= 1 + 1e-9
        for nd in fx["nodes"]:
            assert len(nd) >= 1
            for e in nd:
                assert 0 <= e[0] < npool and len(e) == 1 + k and all(0 <= d < ndir for d in e[1:])
        for g in fx["gpu_first"]:
            if g["status"] == 0:
                assert g["parent_bound"] * (1 - 1e-5) <= g["objective"] and g["lower_bound"] <= g["objective"] * (1 + 1e-7)


def test_oracle_reproduces_a_recorded_config4_bound():
    """CPU-side parity check of the batched engine's recorded output at the config-4 size (200 x 200 block): the exact-eigh oracle
    on the same cut descriptors gives the same bound to 1e-6."""
    from oracle import relaxation as R
    from oracle.cuts import LABELS
    from oracle.datagen import generate_matrix_completion_data
    fx = _load_pool("c4")
    A, mask = generate_matrix_completion_data(fx["k"], fx["n"], fx["m"], fx["n_indices"], fx["seed"])
    q = min((i for i, g in enumerate(fx["gpu_first"]) if g["status"] == 0), key=lambda i: fx["gpu_first"][i]["iters"])
    lab = LABELS[fx["cut_type"]]
    cuts = [(np.array(fx["pool"][e[0]][0]), np.array(fx["pool"][e[0]][1]), [lab[d] for d in e[1:]]) for e in fx["nodes"][q]]
    r = R.solve_relaxation(A, mask, fx["gamma"], fx["k"], fx["cut_type"], cuts, opts=R.Options(eps_abs=1e-8, eps_rel=1e-8, max_iter=6000))
    g = fx["gpu_first"][q]
    assert r["status"] == 0
    assert abs(r["objective"] - g["objective"]) <= 1e-6 * abs(r["objective"]), (r["objective"], g["objective"], r["iters"], g["iters"])
