"""BitMatrix chunk layout and mask compaction (bit-exact integer work).

Julia ``BitMatrix`` layout: linear column-major bit
index ``b = i + n*j`` (0-based) stored in ``chunks[b >> 6]`` bit ``b & 63``, LSB
first.  The reference walks the mask through ``indices[i,j]`` loops
(OMC.jl:2200, 2220, 1852, 2354); the engine compacts it once into row-CSR and
column-CSC.  Pure-Python/NumPy restatement used to check the GPU compaction
bit for bit.  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np


def bitmatrix_chunks(indices):
    """bool (n,m) -> uint64 chunks exactly as ``indices.chunks`` in Julia."""
    n, m = indices.shape
    flat = np.asarray(indices, dtype=bool).T.reshape(-1)  # column-major order
    nchunks = (n * m + 63) // 64
    padded = np.zeros(nchunks * 64, dtype=np.uint8)
    padded[: n * m] = flat
    return np.packbits(padded.reshape(nchunks, 64), axis=1, bitorder="little").view(np.uint64).reshape(-1)


def chunks_to_mask(chunks, n, m):
    bits = np.unpackbits(chunks.view(np.uint8), bitorder="little")[: n * m]
    return bits.reshape((m, n)).T.astype(bool)


def mask_to_csc(indices):
    """Column-compressed: colptr[m+1], rowidx[nnz] ascending inside each column (0-based int32)."""
    n, m = indices.shape
    colptr = np.zeros(m + 1, dtype=np.int32)
    rows = []
    for j in range(m):
        r = np.flatnonzero(indices[:, j]).astype(np.int32)
        rows.append(r)
        colptr[j + 1] = colptr[j] + r.size
    return colptr, (np.concatenate(rows) if rows else np.zeros(0, np.int32))


def mask_to_csr(indices):
    """Row-compressed: rowptr[n+1], colidx[nnz] ascending inside each row (0-based int32)."""
    n, m = indices.shape
    rowptr = np.zeros(n + 1, dtype=np.int32)
    cols = []
    for i in range(n):
        c = np.flatnonzero(indices[i, :]).astype(np.int32)
        cols.append(c)
        rowptr[i + 1] = rowptr[i] + c.size
    return rowptr, (np.concatenate(cols) if cols else np.zeros(0, np.int32))
