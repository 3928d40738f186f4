"""Shared builders for the tests (fixtures -> scipy matrices, random Hi-C-like inputs)."""
import numpy as np
from scipy import sparse


def csr_from_npz(z, prefix, n=None):
    indptr = z[prefix + "_indptr"]
    n = len(indptr) - 1 if n is None else n
    return sparse.csr_matrix((z[prefix + "_data"], z[prefix + "_indices"], indptr), shape=(n, n))


def full_from_pixels(z):
    n = int(z["nbins"])
    up = sparse.csr_matrix((z["count"].astype(np.float64), (z["bin1"], z["bin2"])), shape=(n, n))
    m = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    m.sort_indices()
    return m


def random_hic(n, density, seed, empty_frac=0.0, low_frac=0.0, integer=True):
    """Random symmetric banded-ish count matrix with planted biases."""
    rng = np.random.default_rng(seed)
    nnz_target = int(n * n * density / 2) + n
    i = rng.integers(0, n, nnz_target)
    d = np.minimum(rng.geometric(min(1.0, 8.0 / n), nnz_target) - 1, n - 1)
    j = np.minimum(i + d, n - 1)
    g = np.exp(0.5 * rng.standard_normal(n)).clip(0.2, 5)
    if low_frac:
        g[rng.random(n) < low_frac] = 0.02
    v = 1 + np.floor(40.0 / (1 + np.abs(i - j)) * g[i] * g[j] * rng.random(nnz_target) * 10)
    if not integer:
        v = v * rng.random(nnz_target)
    up = sparse.coo_matrix((v, (i, j)), shape=(n, n)).tocsr()
    up.sum_duplicates()
    up = sparse.triu(up, 0, format="csr")
    m = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    if empty_frac:
        dead = np.flatnonzero(rng.random(n) < empty_frac)
        keep = np.ones(n)
        keep[dead] = 0
        dm = sparse.diags(keep)
        m = sparse.csr_matrix(dm @ m @ dm)
        m.eliminate_zeros()
    m.sort_indices()
    return m.astype(np.float64)
