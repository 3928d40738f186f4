"""numpy restatement of hicTransform's observed/expected methods.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``expected_obs_exp``      <- ``hicexplorer/utilities.py:356-444`` (expected_interactions, single-threaded branch)
* ``expected_non_zero``     <- ``:317-338``
* ``expected_lieberman``    <- ``:293-314``
* ``obs_exp`` / ``obs_exp_non_zero`` / ``obs_exp_lieberman`` <- ``:554-600``, ``:500-551``, ``:479-497`` followed by the
  NaN/Inf scrub + eliminate_zeros of ``hicTransform.py:93-141`` (``_obs_exp`` and friends)
* ``transform``             <- the method / --perChromosome dispatch of ``hicTransform.py:162-215``
* ``covariance`` / ``pearson`` / ``correlation_transform`` <- ``hicTransform.py:105-113, 213-248``: ``np.cov`` /
  ``np.corrcoef`` of the dense block (numpy itself is the dependency that holds the algorithm: ``np.cov`` centres the
  rows, takes ``X @ X.T`` and multiplies by ``1 / (m - 1)``; ``np.corrcoef`` divides by the standard deviations, rows
  then columns, and clips to [-1, 1])

Quirks kept on purpose: the expected value of ``obs_exp`` and ``obs_exp_lieberman`` is looked up at ``ceil(d / 2)``
(``:489, :581``), ``obs_exp`` divides the distance sums by ``n + 1 - d`` (``occurrences``, ``:364``), and the quotient is
cast back to the input's dtype (``.astype(data_type)``), i.e. truncated for integer count matrices.

Pinned (tests/test_oracle.py): equal to the reference's own functions (imported from /root/reference in the build
container) bit for bit on its test matrices and on random input; the golden vectors produced by those functions are
committed in tests/golden/transform_small.npz; the reference's golden files (``test_data/hicTransform/*.cool``) are
matched to the tolerance of the reference's own tests (``decimal=0``; they were written by an older code version).
``correlation_transform(..., "covariance", per_chromosome=True)`` reproduces the reference's golden file
``test_data/hicTransform/covariance_perChromosome.h5`` (same 572 647 upper-triangle cells, values to 1.4e-14;
tests/golden/transform_cov.npz); the reference holds no golden file for ``pearson``, whose vectors in the same npz come
from the reference's own ``_pearson`` run in the build container.
"""
import numpy as np
from scipy import sparse


def _distance_sums(m, sequential=False):
    """sum and count of the non-zero values per distance |row - col|.  ``sequential``: one running sum per distance in
    storage order (the Python loops of utilities.py:301-302, 328-330) instead of ``np.sum`` over each distance's values
    (:424-426) -- the two differ in the last bits for non-integer data."""
    n = m.shape[0]
    coo = m.tocoo()
    keep = coo.data != 0
    row, col, data = coo.row[keep], coo.col[keep], coo.data[keep]
    d = np.abs(row.astype(np.int64) - col.astype(np.int64))
    sums = np.zeros(n)
    if sequential:
        np.add.at(sums, d, data)
        return row, col, data, d, sums, np.bincount(d, minlength=n).astype(np.int64)
    order = np.argsort(d, kind="stable")
    ds, vs = d[order], data[order]
    if len(ds):
        starts = np.flatnonzero(np.r_[True, ds[1:] != ds[:-1]])
        ends = np.r_[starts[1:], len(ds)]
        for a, b in zip(starts, ends):
            sums[ds[a]] = np.sum(vs[a:b])              # np.sum(pSubmatrix.data[mask]) of utilities.py:424-426
    counts = np.bincount(d, minlength=n).astype(np.int64)
    return row, col, data, d, sums, counts


def _scrub(x):
    x[np.isnan(x)] = 0
    x[np.isinf(x)] = 0
    return x


def expected_obs_exp(sums, n):
    with np.errstate(divide="ignore", invalid="ignore"):
        return _scrub(np.divide(sums, np.arange(n + 1, 1, -1)))


def expected_non_zero(sums, counts):
    with np.errstate(divide="ignore", invalid="ignore"):
        return _scrub(np.divide(sums, counts.astype(np.float64)))


def expected_lieberman(sums, length_chromosome, chromosome_count):
    count_times_i = np.arange(float(len(sums)))
    count_times_i *= np.int32(chromosome_count)
    count_times_i -= np.int32(length_chromosome)
    count_times_i *= np.int32(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.divide(sums, count_times_i)


def _finish(values, row, col, shape):
    out = sparse.csr_matrix((values, (row, col)), shape=shape)
    out.data = _scrub(out.data) if out.data.dtype.kind == "f" else out.data
    out.eliminate_zeros()
    out.sort_indices()
    return out


def _half_index_quotient(m, expected, row, col, data, d):
    idx = np.ceil(d / 2).astype(np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.divide(data.astype(np.float32), expected[idx])
        q = _scrub(q)
        return q.astype(m.dtype)


def obs_exp(m):
    if m.nnz == 0:
        return sparse.csr_matrix(m)
    row, col, data, d, sums, _counts = _distance_sums(m)
    if len(d) == 0:
        return sparse.csr_matrix(m.shape, dtype=m.dtype)
    return _finish(_half_index_quotient(m, expected_obs_exp(sums, m.shape[0]), row, col, data, d), row, col, m.shape)


def obs_exp_lieberman(m, length_chromosome, chromosome_count):
    if m.nnz == 0:
        return sparse.csr_matrix(m)
    row, col, data, d, sums, _counts = _distance_sums(m, sequential=True)
    expected = expected_lieberman(sums, length_chromosome, chromosome_count)
    return _finish(_half_index_quotient(m, expected, row, col, data, d), row, col, m.shape)


def obs_exp_non_zero(m, ligation_factor=False):
    if m.nnz == 0:
        return sparse.csr_matrix(m)
    row, col, data, d, sums, counts = _distance_sums(m, sequential=True)
    expected = expected_non_zero(sums, counts)[d]
    if ligation_factor:
        row_sums = np.array(m.sum(axis=1).T).flatten()
        total = m.sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            expected = expected * (row_sums[row] * row_sums[col] / total)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        q = np.divide(data.astype(np.float32).astype(np.float64), expected).astype(np.float32)
    return _finish(q, row, col, m.shape)


def transform(matrix, method, chr_offsets=None, per_chromosome=False, ligation_factor=False):
    """the dispatch of hicTransform.main for the three obs/exp methods; ``matrix`` = full symmetric CSR"""
    m = sparse.csr_matrix(matrix)
    if method == "obs_exp_lieberman" or per_chromosome:
        offs = np.asarray(chr_offsets, dtype=np.int64)
        total_len, n_chr = int(offs[-1] - offs[0]), len(offs) - 1
        blocks = []
        for lo, hi in zip(offs[:-1], offs[1:]):
            sub = sparse.csr_matrix(m[lo:hi, lo:hi])
            if method == "obs_exp":
                blocks.append(obs_exp(sub))
            elif method == "obs_exp_non_zero":
                blocks.append(obs_exp_non_zero(sub, ligation_factor))
            else:
                blocks.append(obs_exp_lieberman(sub, total_len, n_chr))
        out = sparse.block_diag(blocks, format="csr").astype(np.float64)   # assembled in a float64 lil_matrix (:161)
        out.sort_indices()
        return out
    if method == "obs_exp":
        return obs_exp(m)
    if method == "obs_exp_non_zero":
        return obs_exp_non_zero(m, ligation_factor)
    raise ValueError(method)


def covariance(dense):
    """hicTransform.py:236, 243: np.cov of the dense block, every stored NaN kept"""
    return sparse.csr_matrix(np.cov(np.asarray(dense, dtype=np.float64)))


def pearson(dense):
    """hicTransform.py:105-113 (_pearson): np.corrcoef, NaN / Inf -> 0"""
    dense = np.asarray(dense)
    if dense.shape[0] == 0:
        return sparse.csr_matrix(dense.shape, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.corrcoef(dense)
    c = sparse.csr_matrix(c)
    c.data[np.isnan(c.data)] = 0
    c.data[np.isinf(c.data)] = 0
    return c


def correlation_transform(matrix, method, chr_offsets=None, per_chromosome=False):
    """the pearson / covariance branches of hicTransform.main (:213-248); zeros eliminated like the saved file"""
    m = sparse.csr_matrix(matrix)
    fn = pearson if method == "pearson" else covariance
    if per_chromosome:
        offs = np.asarray(chr_offsets, dtype=np.int64)
        out = sparse.block_diag([fn(m[lo:hi, lo:hi].toarray()) for lo, hi in zip(offs[:-1], offs[1:])], format="csr")
    else:
        out = sparse.csr_matrix(fn(m.toarray()))
    out = out.astype(np.float64)
    out.eliminate_zeros()
    out.sort_indices()
    return out
