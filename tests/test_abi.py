"""CPU tests of the drop-in boundary: libomc_b200.so loads, exports every symbol include/omc_b200.h declares,
and fails loudly (no CPU fallback) when no GPU is usable.  No compute calls are made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "omc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(omc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_exported_and_bound():
    import omc_b200
    from omc_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/omc_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes and header disagree"


def test_opts_struct_layout_and_defaults():
    import omc_b200
    o = omc_b200.default_opts()
    assert ctypes.sizeof(o) == 88
    assert o.eps_abs == 1e-8 and o.eps_rel == 1e-8 and o.max_iter == 20000
    assert o.check_every == 25 and o.adapt_every == 100 and o.fix_linear3_right == 0
    assert o.rho0 == 0.3 and o.sigma == 1e-6 and o.alpha == 1.6 and o.cutoff == float("inf")
    with pytest.raises(TypeError):
        omc_b200.default_opts(nonsense=1)


def test_compute_entry_fails_loudly_without_init_or_gpu():
    """The product path never falls back to the CPU: without a device omc_init reports an error and
    every compute entry refuses to run."""
    import torch
    from omc_b200 import _lib
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    rc = lib.omc_init(0)
    assert rc == -2 and b"no CPU fallback" in lib.omc_last_error()
    out = (ctypes.c_double * 2)()
    assert lib.omc_measure_fp64_peak(out) == -3        # OMC_ERR_STATE: not initialised
    assert b"omc_init" in lib.omc_last_error()


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "optimalmatrixcompletion.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"


def test_bitmatrix_chunks_binding_matches_oracle_layout():
    import numpy as np
    import omc_b200
    from oracle.mask import bitmatrix_chunks
    rng = np.random.default_rng(0)
    for n, m in [(3, 5), (10, 10), (50, 50), (7, 64)]:
        mask = rng.random((n, m)) < 0.5
        assert np.array_equal(omc_b200.bitmatrix_chunks(mask), bitmatrix_chunks(mask))
