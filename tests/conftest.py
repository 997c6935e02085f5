import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "relax_golden.json")) as f:
        return json.load(f)


def golden_instance(case):
    """Rebuilds (A, mask) of a golden case from its seed with the committed generator."""
    from oracle.datagen import generate_matrix_completion_data
    A, mask = generate_matrix_completion_data(case["k"], case["n"], case["m"], case["n_indices"], case["seed"])
    cuts = [(np.array(c["x"]), np.array(c["Uhat"]), list(c["dirs"])) for c in case["cuts"]]
    return A, mask, cuts


def cut_region(ct, v, h):
    a = abs(h)
    if ct == "linear":
        return "left" if v <= h else "right"
    if ct == "linear2":
        return "left" if v <= -a else ("middle" if v <= a else "right")
    return "left" if v <= -a else ("inner_left" if v <= 0 else ("inner_right" if v <= a else "right"))


def feasible_chain(ct, n, k, L, rng):
    """L cuts whose regions all contain one hidden rank-k factor (orthonormal columns, bottom k x k block a positive
    diagonal so that the sign normalisation OMC.jl:1442-1449 holds)."""
    W, _ = np.linalg.qr(rng.standard_normal((n - k, k)))
    th = rng.uniform(0.3, 1.2, size=k)
    Us = np.vstack([W * np.cos(th), np.diag(np.sin(th))])
    out = []
    for _ in range(L):
        x = rng.standard_normal(n); x /= np.linalg.norm(x)
        Uh = rng.uniform(-0.4, 0.4, size=(1, k)) * x[:, None]
        out.append((x, Uh, [cut_region(ct, float(x @ Us[:, j]), float(Uh[:, j] @ x)) for j in range(k)]))
    return out
