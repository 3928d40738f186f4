"""Device-resident correction pipelines: what ``hicCorrectMatrix correct`` does
between loading the matrix and saving it (hicexplorer/hicCorrectMatrix.py:612-757),
with the matrix uploaded ONCE and every stage (zero-bin mask, NaN/Inf scrub,
MAD filter, bin deletion, ICE / KR) running on the GPU.

Host code here only sequences C-ABI calls and does O(n) index bookkeeping.
"""
import logging

import numpy as np
from scipy import sparse

from . import _lib
from .store import KR_DIAG_EPS, DeviceCSR, canonical_csr

log = logging.getLogger(__name__)


def _segment_offsets_after(mask_removed, offsets):
    """chromosome offsets after deleting the bins flagged in ``mask_removed``."""
    kept_before = np.concatenate([[0], np.cumsum(~mask_removed.astype(bool))])
    return kept_before[np.asarray(offsets, dtype=np.int64)]


def ice_pipeline(matrix, filter_threshold, max_iter=500, tolerance=1e-5, chr_offsets=None, perchr=False,
                 skip_diagonal=False, device=0, want_matrix=True, extra_mask_fn=None):
    """ICE branch of the driver.

    matrix           full symmetric scipy matrix in the original bin frame (n bins)
    filter_threshold (lo, hi) modified z-score thresholds (--filterThreshold)
    chr_offsets      bin offsets of the chromosomes (len nchrom+1); needed for perchr
    perchr           statistics and correction per chromosome block (--perchr)
    extra_mask_fn    optional callable(kept_bins) -> bool mask over the bins that survived the MAD
                     filter, for the host-side --sequencedCountCutoff filter

    Returns a dict:
      zero_bins      original ids removed because their row sum is 0        (:612-616)
      outliers_rel   MAD outliers, indices relative to the zero-masked frame (:656-657)
      kept           original ids of the bins that went through ICE
      bias           correction factors of the kept bins                     (:738 / :713)
      matrix         corrected CSR in the kept frame (block diagonal for perchr)
      iterations     loop bodies executed (per chromosome for perchr)
      deviation      last max|s-1| (per chromosome for perchr)
    """
    _lib.require_device()
    m = canonical_csr(matrix)
    n = m.shape[0]
    lo, hi = filter_threshold
    with DeviceCSR.from_scipy(m, device=device) as dev:
        # mask all zero value bins: row sums include the diagonal and NaNs propagate, as in the
        # reference where this happens before the NaN/Inf scrub
        row_sum, _ = dev.row_stats()
        zero_mask = row_sum == 0
        zero_bins = np.flatnonzero(zero_mask)
        log.info("Removing {} zero value bins".format(len(zero_bins)))
        offsets = None if chr_offsets is None else np.asarray(chr_offsets, dtype=np.int64)
        if perchr:
            if offsets is None:
                raise ValueError("perchr needs chromosome offsets")
            dev.set_segments(offsets)
        dev.compact(zero_mask.astype(np.uint8))
        keep1 = np.flatnonzero(~zero_mask)
        dev.scrub()
        log.info("matrix contains {} data points. Sparsity {:.3f}.".format(
            dev.nnz, float(dev.nnz) / (max(dev.n, 1) ** 2)))
        if skip_diagonal:
            dev.diag_zero()
        mask, _z, _med, _mad = dev.mad_filter(lo, hi)
        outliers_rel = np.flatnonzero(mask)
        remove = mask.astype(bool)
        if extra_mask_fn is not None:
            extra = np.zeros(len(keep1), dtype=bool)
            surv = np.flatnonzero(~remove)
            extra[surv] = extra_mask_fn(keep1[surv])
            remove |= extra
        dev.compact(remove.astype(np.uint8))
        kept = keep1[~remove]
        if perchr:
            dev.keep_cis()
        if dev.symmetry_ratio() > 1e-10:
            raise ValueError("Please provide symmetric matrix!")
        bias, iters, deviation = dev.ice(max_iter, tolerance)
        out = dict(zero_bins=zero_bins, outliers_rel=outliers_rel, kept=kept, bias=bias,
                   iterations=iters, deviation=deviation, n=n)
        if want_matrix:
            values = dev.ice_scaled_values()
            pattern = dev.download()
            out["matrix"] = sparse.csr_matrix((values, pattern.indices, pattern.indptr), shape=pattern.shape)
    return out


def kr_pipeline(matrix, rescale=True, want_matrix=True, device=0):
    """KR branch of the driver (hicCorrectMatrix.py:742-757): no masking, no filter.
    Returns dict(x, factor, outer, matvecs, residual, upper)."""
    _lib.require_device()
    m = canonical_csr(matrix)
    with DeviceCSR.from_scipy(m, device=device) as dev:
        x, outer, matvecs, residual = dev.kr()
        factor = 1.0
        if rescale:
            x, factor = dev.kr_rescale(x, KR_DIAG_EPS)
        out = dict(x=x, factor=factor, outer=outer, matvecs=matvecs, residual=residual)
        if want_matrix:
            out["upper"] = dev.kr_scaled_upper(x, KR_DIAG_EPS)
    return out


def scatter_to_frame(corrected_kept, kept, n):
    """Re-insert the masked bins as empty rows/columns (hicmatrix restoreMaskedBins)."""
    coo = corrected_kept.tocoo()
    kept = np.asarray(kept)
    return sparse.csr_matrix((coo.data, (kept[coo.row], kept[coo.col])), shape=(n, n))
