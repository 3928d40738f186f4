"""numpy restatement of the reference's ICE path.  TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``): the product never imports this module.

What is restated, and from where (paths relative to ``/root/reference``):

* ``ice()``            <- ``hicexplorer/iterativeCorrection.py:10-86``
* ``MadScores``        <- ``hicexplorer/hicCorrectMatrix.py:349-391`` (class MAD)
* ``outlier_bins()``   <- ``hicexplorer/hicCorrectMatrix.py:548-592`` (filter_by_zscore)
* ``zero_bins()``      <- ``hicexplorer/hicCorrectMatrix.py:612-617``
* ``scrub()``          <- ``hicexplorer/utilities.py:111-128`` + ``hicCorrectMatrix.py:624-626``
* ``symmetrize_triu`` / ``select_bins`` / ``correct_ice_pipeline`` restate
  the few ``hicmatrix`` operations the driver calls;
  hicmatrix itself is a third-party package that is not under
  ``/root/reference``.

Pinned (tests/test_oracle.py): against the reference's own golden file
``test_data/hicCorrectMatrix/small_test_matrix_ICEcorrected_chrUextra_chr3LHet.h5``
bit-exactly (data, indices, indptr, correction_factors, nan_bins), against
``filtered.bed`` (the 64 MAD outliers), and -- in the build container, where
``/root/reference`` is importable -- against the reference functions
themselves on random inputs.

The arithmetic deliberately keeps the reference's operation ORDER (two
separate in-place multiplies per iteration, reciprocal then multiply, COO row
sums) so that results are bit-identical to the reference, not merely close.
"""
import numpy as np
from scipy import sparse


class IceOverflow(Exception):
    """The reference calls ``exit(1)`` here (iterativeCorrection.py:58-62, 80-84)."""

    def __init__(self, stage):
        Exception.__init__(self, "ICE overflow (%s)" % stage)
        self.stage = stage


def is_symmetric(csr):
    """iterativeCorrection.py:35 -- mean|M - M^T| / |mean M| <= 1e-10."""
    num = np.abs(csr - csr.T).mean()
    den = 1.0 * np.abs(csr.mean())
    return not (num / den > 1e-10)


def ice(matrix, max_iter=50, tolerance=1e-5):
    """Literal ICE.  Returns dict(matrix=CSR float64, bias, iterations, deviation).

    ``iterations`` is the number of loop bodies executed (the reference logs
    this value + 1, iterativeCorrection.py:41,73).
    """
    n = matrix.shape[0]
    bias = np.ones(n, dtype=np.float64)
    if np.isnan(matrix.sum()):
        matrix.data[np.isnan(matrix.data)] = 0  # mutates the caller's matrix, like the reference
    work = matrix.astype(float)
    coo = work.tocoo()
    if not is_symmetric(work):
        raise ValueError("Please provide symmetric matrix!")
    rows, cols, vals = coo.row, coo.col, coo.data
    done = 0
    deviation = np.nan
    for _ in range(max_iter):
        done += 1
        marg = np.array(coo.sum(axis=1)).flatten()
        live = marg != 0
        marg = marg / np.mean(marg[live])
        bias *= marg
        deviation = np.abs(marg - 1).max()
        inv = 1.0 / marg
        vals *= np.take(inv, rows)
        vals *= np.take(inv, cols)
        if np.any(vals > 1e100):
            raise IceOverflow("iteration")
        if deviation < tolerance:
            break
    scale = bias[bias != 0].mean()
    bias = np.divide(bias, scale)
    coo.data = vals * scale * scale
    if np.any(coo.data > 1e10):
        raise IceOverflow("final")
    return dict(matrix=coo.tocsr(), bias=bias, iterations=done, deviation=float(deviation))


class MadScores(object):
    """Modified z-score (Iglewicz & Hoaglin) exactly as the reference's MAD class."""

    B = 0.6745

    def __init__(self, points):
        points = np.asarray(points)
        col = points[:, None] if points.ndim == 1 else points
        self.median = np.median(col[col > 0], axis=0)
        diff = np.sum(col - self.median, axis=-1)
        self.mad = np.median(np.abs(diff))
        self.z = np.multiply(self.B, np.divide(diff, self.mad))

    def outliers(self, lo, hi):
        return (self.z < lo) | (self.z > hi)


def coverage_points(csr):
    """Row sum minus diagonal (hicCorrectMatrix.py:584-588)."""
    return np.asarray(csr.sum(axis=1)).flatten() - csr.diagonal()


def outlier_bins(csr, lo, hi, chr_ranges=None):
    """Sorted bin ids whose modified z-score is outside (lo, hi).

    ``chr_ranges`` = list of (start, end) row ranges -> the ``--perchr`` variant
    (hicCorrectMatrix.py:557-581): statistics per diagonal block.
    """
    if chr_ranges is None:
        return np.flatnonzero(MadScores(coverage_points(csr)).outliers(lo, hi))
    found = []
    for a, b in chr_ranges:
        block = csr[a:b, a:b]
        block.data[np.isnan(block.data)] = 0
        ids = np.flatnonzero(MadScores(coverage_points(block)).outliers(lo, hi))
        found.extend((ids + a).tolist())
    return np.array(sorted(found), dtype=np.int64)


def diagnostic_histograms(csr, chr_ranges=None, x_max=None):
    """The numbers behind ``hicCorrectMatrix diagnostic_plot`` (hicCorrectMatrix.py:425-545, the matplotlib calls left
    out): per histogram (one for the whole matrix, or one per chromosome block with ``chr_ranges``) the coverage values
    with a modified z-score below 5, their 100-bin histogram (``ax.hist(values, 100)`` = ``np.histogram``), and the
    modified z-score of the first local minimum of the histogram -- the "mad threshold" the tool logs and marks."""
    # main() masks the zero-value bins for the plot as well (hicCorrectMatrix.py:616-619), on row sums taken before
    # the NaN / Inf scrub; the chromosome ranges shrink accordingly
    rs = np.asarray(csr.sum(axis=1)).flatten()
    keep = rs != 0
    if not keep.all():
        before = np.concatenate([[0], np.cumsum(keep)])
        csr = select_bins(csr, np.flatnonzero(keep))
        if chr_ranges is not None:
            chr_ranges = [(int(before[a]), int(before[b])) for a, b in chr_ranges]
    m = scrub(csr)
    out = []
    for a, b in (chr_ranges if chr_ranges is not None else [(0, m.shape[0])]):
        block = m[a:b, a:b] if chr_ranges is not None else m
        points = coverage_points(block)
        mad = MadScores(points)
        values = points[mad.z < 5]
        if x_max:
            values = values[values < x_max]
        if len(values) == 0:
            out.append(dict(values=values, counts=None, edges=None, threshold=None, mad_threshold=None))
            continue
        counts, edges = np.histogram(values, 100)
        local_min = [x for x, y in enumerate(counts) if 1 <= x < len(counts) - 1 and counts[x - 1] > y < counts[x + 1]]
        threshold = edges[local_min[0]] if local_min else None
        mad_threshold = None
        if threshold:
            diff = threshold - mad.median
            mad_threshold = mad.B * diff if mad.mad == 0.0 else mad.B * diff / mad.mad
        out.append(dict(values=values, counts=counts, edges=edges, threshold=threshold, mad_threshold=mad_threshold))
    return out


def zero_bins(csr):
    """Bins whose row sum (diagonal included) is exactly 0 (hicCorrectMatrix.py:612-616)."""
    return np.flatnonzero(np.asarray(csr.sum(axis=1)).flatten() == 0)


def scrub(csr):
    """NaN -> 0, Inf -> 0, cast to float64 (hicCorrectMatrix.py:624-626)."""
    csr = csr.copy()
    csr.data[np.isnan(csr.data)] = 0
    csr.data[np.isinf(csr.data)] = 0
    return csr.astype(np.float64, copy=True)


# ---- the hicmatrix operations the driver relies on ----
def symmetrize_triu(triu_csr):
    """File holds the upper triangle; in memory the matrix is full symmetric."""
    triu_csr = sparse.csr_matrix(triu_csr)
    return sparse.csr_matrix(triu_csr + sparse.triu(triu_csr, 1).T)


def select_bins(csr, keep):
    keep = np.asarray(keep)
    return sparse.csr_matrix(csr[keep, :][:, keep])


def chromosome_order(chr_list, wanted):
    """Bin ids of the chromosomes in ``wanted``, in that order (reorderChromosomes)."""
    chr_list = np.asarray(chr_list)
    parts = [np.flatnonzero(chr_list == c) for c in wanted]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)


def trunc_trans(csr, chr_offsets, high, clip):
    """hicmatrix ``truncTrans(high)`` as called at hicCorrectMatrix.py:695-698 (hicmatrix is third party and not under
    /root/reference; restated from its published source [memory]): ``max_inter = np.percentile(trans values,
    100 - high)``, trans = row and column in different chromosomes.  The documented effect is to clip the trans values
    ``>= max_inter`` to ``max_inter`` (``clip=True``); the published statement reads
    ``mat.data[...] == max_inter`` -- a comparison, i.e. no change (``clip=False``).  Returns (matrix, max_inter)."""
    coo = csr.tocoo()
    offs = np.asarray(chr_offsets)
    trans = np.searchsorted(offs, coo.row, side="right") != np.searchsorted(offs, coo.col, side="right")
    if np.count_nonzero(trans) == 0:
        return csr, None
    max_inter = np.percentile(coo.data[trans], (100 - high))
    if clip:
        data = coo.data.copy()
        data[(data >= max_inter) & trans] = max_inter
        csr = sparse.csr_matrix((data, (coo.row, coo.col)), shape=csr.shape)
    return csr, max_inter


def inflated_bins(corrected_csr, pre_row_sum, cutoff):
    """hicCorrectMatrix.py:767-771: bins whose row sum grew by at least ``cutoff`` times."""
    after = np.asarray(corrected_csr.sum(axis=1)).flatten()
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.flatnonzero(after / pre_row_sum >= cutoff)


def correct_ice_pipeline(full_csr, lo, hi, max_iter=500, tolerance=1e-5, chr_offsets=None, trans_cutoff=None,
                         clip_trans=False, inflation_cutoff=None):
    """zero-bin mask -> scrub -> MAD mask -> [truncTrans] -> ICE -> [inflation mask] -> scatter back into the
    original frame; what ``hicCorrectMatrix correct --correctionMethod ICE``
    does for a genome-wide run (hicCorrectMatrix.py:612-673, 695-699, 735-740, 759-779).

    Returns dict with the stages so every one of them can be compared.
    """
    n = full_csr.shape[0]
    dead = zero_bins(full_csr)
    keep1 = np.setdiff1d(np.arange(n), dead)
    stage1 = scrub(select_bins(full_csr, keep1))
    out_rel = outlier_bins(stage1, lo, hi)
    keep2 = np.delete(np.arange(stage1.shape[0]), out_rel)
    ice_in = select_bins(stage1, keep2)
    orig = keep1[keep2]
    extra = {}
    pre_row_sum = None
    if trans_cutoff and 0 < trans_cutoff < 100:
        offs_now = np.searchsorted(orig, np.asarray(chr_offsets), side="left")     # offsets in the masked frame
        ice_in, extra["max_inter"] = trunc_trans(ice_in, offs_now, float(trans_cutoff) / 100, clip_trans)
        pre_row_sum = np.asarray(ice_in.sum(axis=1)).flatten()
        extra["pre_row_sum"] = pre_row_sum
    res = ice(ice_in.copy(), max_iter=max_iter, tolerance=tolerance)
    if inflation_cutoff and inflation_cutoff > 0:
        infl = inflated_bins(res["matrix"], pre_row_sum, inflation_cutoff)     # NameError in the reference without transCutoff
        extra["inflated_rel"] = infl
        keep3 = np.delete(np.arange(len(orig)), infl)
        res = dict(res, matrix=select_bins(res["matrix"].tocsr(), keep3), bias=res["bias"][keep3])
        orig = orig[keep3]
    coo = res["matrix"].tocoo()
    corrected_full = sparse.coo_matrix((coo.data, (orig[coo.row], orig[coo.col])), shape=(n, n)).tocsr()
    upper = sparse.triu(corrected_full, 0, format="csr")
    upper.eliminate_zeros()
    upper.sort_indices()
    factors = np.full(n, np.nan)
    factors[orig] = res["bias"]
    nan_bins = np.setdiff1d(np.arange(n), orig).astype(np.int64)
    return dict(zero_bins=dead, outliers_rel=out_rel, outliers_abs=keep1[out_rel], kept=orig,
                ice_input=ice_in, ice=res, upper=upper, correction_factors=factors, nan_bins=nan_bins, **extra)
