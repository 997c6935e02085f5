"""CPU oracle for the OptimalMatrixCompletion.jl bounding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker / reported
CPU baseline.  The product path (``optimalmatrixcompletion.jl_b200``) never
imports this package and fails loudly when its CUDA library is missing.

PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, fixtures or
golden vectors (``/root/reference/test/runtests.jl`` is empty) and its
arithmetic lives in un-vendored third-party binaries (Mosek 10.1.1, ARPACK
3.5.1, OpenBLAS 0.3.21 -- ``/root/reference/Manifest.toml``), none of which
exist in this image; neither does Julia.  The oracle is therefore a NumPy/SciPy
restatement of the reference's *mathematical program* (file:line cited on every
function), pinned instead by
  * analytic known-answer tests (``oracle/kat.py``),
  * solver-independent optimality certificates (primal feasibility + dual
    feasibility + zero gap, ``oracle/relaxation.py:certificate``),
  * an independent second algorithm for the root node
    (``oracle/kat.py:root_bound_projected_gradient``).

``OMC.jl:N`` below means ``/root/reference/src/OptimalMatrixCompletion.jl:N``.
"""
