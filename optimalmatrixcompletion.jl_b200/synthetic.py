"""Synthetic benchmark inputs with the distribution of the reference's generators (host side, NumPy).

Restates /root/reference/src/utils.jl:3-26 (mask with exactly ``n_indices`` observed entries, redrawn until every row and
column holds one, at most 100 redraws) and utils.jl:97-103 (``A = L R + 0.01 E`` with i.i.d. Gaussian factors).  Julia's
MersenneTwister streams cannot be reproduced without Julia, so the seeds are NumPy ``default_rng`` seeds.
This is what ``bench.py`` feeds the engine; the oracle keeps its own copy so that the product never imports ``oracle/``.
"""
import numpy as np


def generate_masked_bitmatrix(n, m, sparsity, rng, max_iters=100):
    it = 0
    while True:
        flat = np.zeros(n * m, dtype=bool)
        flat[rng.permutation(n * m)[:sparsity]] = True
        indices = flat.reshape((m, n)).T.copy()          # Julia's reshape is column-major: linear index b = i + n j
        if (indices.any(axis=0).all() and indices.any(axis=1).all()) or it >= max_iters:
            return indices
        it += 1


def generate_matrix_completion_data(k, n, m, n_indices, seed, eps=0.01):
    if not n <= m:
        raise ValueError("Input matrix A must have size (n, m) with n <= m.")
    if n_indices > n * m:
        raise ValueError("n_indices exceeds n*m")
    rng = np.random.default_rng(seed)
    left = rng.standard_normal((n, k))
    right = rng.standard_normal((k, m))
    noise = rng.standard_normal((n, m))
    A = left @ right + eps * noise
    indices = generate_masked_bitmatrix(n, m, n_indices, rng)
    return np.asfortranarray(A), np.asfortranarray(indices)
