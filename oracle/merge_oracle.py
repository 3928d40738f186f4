"""numpy restatement of bin merging (``hicMergeMatrixBins --numBins k``), the step right before correction in the
reference's documented workflow.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restated from (paths relative to ``/root/reference``):
* ``merge_groups()``   <- ``hicexplorer/hicMergeMatrixBins.py:228-262`` (which bins are merged, new intervals)
* ``reduce()``         <- ``hicexplorer/reduceMatrix.py:12, 141-232`` with ``use_triu=True, diagonal=True``

Pinned (tests/test_oracle.py) against the reference's golden file
``test_data/hicMergeMatrixBins/result.h5`` (``small_test_matrix.h5 --numBins 5``; the reference's own test compares
``matrix.data`` and ``cut_intervals`` exactly, test/general/test_hicMergeMatrixBins.py:15-27) and, in the build
container, against ``reduce_matrix`` itself on random inputs.
"""
import numpy as np
from scipy import sparse


def merge_groups(intervals, num_bins):
    """intervals: list of (chrom, start, end, coverage).  Returns (bin_map int32[n] with -1 for dropped bins,
    new_intervals).  Consecutive bins of a chromosome are grouped ``num_bins`` at a time; a trailing group of a
    chromosome with fewer than ``num_bins / 2`` bins is dropped, except the very last group of the matrix."""
    chrom, start, end, cov = zip(*intervals)
    n = len(chrom)
    bin_map = np.full(n, -1, dtype=np.int32)
    new_intervals = []
    first = 0
    count = 0
    prev = chrom[0]
    for idx in range(n):
        if (count > 0 and count % num_bins == 0) or chrom[idx] != prev:
            if not count < num_bins / 2:
                bin_map[first:idx] = len(new_intervals)
                new_intervals.append((chrom[first], start[first], end[idx - 1], float(np.mean(cov[first:idx]))))
            first = idx
            count = 0
        prev = chrom[idx]
        count += 1
    bin_map[first:n] = len(new_intervals)
    new_intervals.append((chrom[n - 1], start[first], end[n - 1], float(np.mean(cov[first:]))))
    return bin_map, new_intervals


def reduce(matrix, bin_map, n_new):
    """Sum the upper triangle into the merged bins, then mirror it (the diagonal is kept once)."""
    up = sparse.triu(matrix, k=0, format="coo")
    r = bin_map[up.row]
    c = bin_map[up.col]
    keep = (r > -1) & (c > -1)
    acc = sparse.coo_matrix((up.data[keep], (r[keep], c[keep])), shape=(n_new, n_new)).tocsr()   # sums duplicates
    diag = sparse.dia_matrix(([acc.diagonal()], [0]), shape=(n_new, n_new))
    full = sparse.csr_matrix(acc + acc.T - diag)
    full.eliminate_zeros()
    full.sort_indices()
    return full


def merge_bins(matrix, intervals, num_bins):
    bin_map, new_intervals = merge_groups(intervals, num_bins)
    merged = reduce(matrix, bin_map, len(new_intervals))
    nan_bins = np.flatnonzero(np.asarray(merged.sum(axis=0)).ravel() == 0)
    return merged, new_intervals, nan_bins, bin_map


def running_window(matrix, num_bins):
    """``running_window_merge`` (hicexplorer/hicMergeMatrixBins.py:88-186) restated: the reference concatenates the
    num_bins^2 shifted copies of the upper triangle's COO triplets and drops the ones that leave the frame -- that is
    S U S^T with the banded 0/1 matrix S[a, i] = (|a - i| <= h) -- then keeps triu and mirrors it."""
    if num_bins == 1:
        return sparse.csr_matrix(matrix)
    assert num_bins % 2 == 1, "num_bins has to be an odd number"
    half = (num_bins - 1) // 2
    n = matrix.shape[0]
    offsets = list(range(-half, half + 1))
    band = sparse.diags([np.ones(n - abs(d), dtype=matrix.dtype) for d in offsets], offsets, format="csr", dtype=matrix.dtype)
    upper = sparse.triu(matrix, 0, format="csr")
    new = sparse.triu(band @ upper @ band.T, 0, format="csr")
    full = sparse.csr_matrix(new + new.T - sparse.diags(new.diagonal(), dtype=matrix.dtype)).astype(matrix.dtype)
    full.eliminate_zeros()
    full.sort_indices()
    return full
