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
