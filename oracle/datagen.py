"""Synthetic instances with the distribution of the reference's generators.

Follows ``/root/reference/src/utils.jl:3-26`` (mask with exactly ``n_indices``
entries, redrawn until every row and column is hit, at most ``max_iters``
redraws) and ``utils.jl:97-103`` (``A = L R + eps * E``, Gaussian factors).
Julia's MersenneTwister streams cannot be reproduced without Julia, so the
seeds are NumPy ``default_rng`` seeds; only the
*distribution* is the reference's.
"""
import numpy as np

# (k, n, m, n_indices) of the five BASELINE.json configs
CONFIGS = {
    "C1": dict(k=1, n=10, m=10, n_indices=50, cut_type="linear", nev=1),
    "C2": dict(k=1, n=50, m=50, n_indices=1250, cut_type="linear", nev=1),
    "C3": dict(k=2, n=30, m=30, n_indices=450, cut_type="linear2", nev=1),
    "C4": dict(k=3, n=100, m=100, n_indices=3000, cut_type="linear3", nev=2),
    "C5": dict(k=5, n=1000, m=1000, n_indices=200000, cut_type="linear", nev=1),
}
GAMMA = 80.0  # README.md:33


def generate_masked_bitmatrix(n, m, sparsity, rng, max_iters=100):
    """utils.jl:3-26 -- exactly ``sparsity`` observed entries, all rows/cols hit."""
    it = 0
    while True:
        flat = np.zeros(n * m, dtype=bool)
        flat[rng.permutation(n * m)[:sparsity]] = True
        # Julia reshape is column-major: linear index b = i + n*j
        indices = flat.reshape((m, n)).T.copy()
        if (indices.any(axis=0).all() and indices.any(axis=1).all()) or it >= max_iters:
            return indices
        it += 1


def generate_matrix_completion_data(k, n, m, n_indices, seed, eps=0.01):
    """utils.jl:69-110 -- low-rank Gaussian product plus ``eps`` Gaussian noise."""
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


def config_instance(name, seed=0):
    c = CONFIGS[name]
    A, mask = generate_matrix_completion_data(c["k"], c["n"], c["m"], c["n_indices"], seed)
    return c["k"], A, mask, GAMMA
