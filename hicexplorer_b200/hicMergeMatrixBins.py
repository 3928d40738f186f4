"""``hicMergeMatrixBins`` on a B200: same flags and output as the reference tool
(hicexplorer/hicMergeMatrixBins.py:16-283), the summation on the GPU (``hb_csr_merge_bins``).

This is the step right before correction in the reference's documented workflow.
``--numBins`` up to 1024 (more than 32 bins per group are merged in two exact device steps).  ``--runningWindow``
(hicMergeMatrixBins.py:88-186: every cell becomes the sum of the (numBins x numBins) box around it, taken over the upper
triangle, resolution unchanged) runs on the device as the sum of the numBins^2 shifted copies of the upper triangle
(``hb_csr_shifted_copy`` + ``hb_csr_add``), cut back to the upper triangle and mirrored (``hb_csr_symmetrize_upper``).
"""
import argparse
import logging

import numpy as np

from . import hicmatrix_lite as hm
from ._version import __version__
from .store import DeviceCSR

log = logging.getLogger(__name__)


def parse_arguments(args=None):
    parser = argparse.ArgumentParser(
        formatter_class=argparse.ArgumentDefaultsHelpFormatter, add_help=False,
        description='Merges bins from a Hi-C matrix. For example, using a matrix containing 5kb bins, a matrix '
                    'of 50kb bins can be derived using --numBins 10.')
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrix', '-m', help='Matrix to reduce in h5 format.', metavar='.h5 file format', required=True)
    req.add_argument('--outFileName', '-o', help='File name to save the resulting matrix. The output is also a .h5 '
                     'file. But don\'t add the suffix.', required=True)
    req.add_argument('--numBins', '-nb', help='Number of bins to merge.', metavar='int', type=int, required=True)
    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--runningWindow', help='set to merge for using a running window of length --numBins. Must be '
                     'an odd number.', action='store_true')
    opt.add_argument('--help', '-h', action='help', help='show this help message and exit')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    opt.add_argument('--device', help='CUDA device index.', type=int, default=0)
    return parser


def remove_nans_if_needed(hic):
    """corrected matrices carry NaN bins; they are removed before merging (hicMergeMatrixBins.py:66-85)"""
    if hic.nan_bins is not None and len(hic.nan_bins):
        hic.maskBins(hic.nan_bins)
        hic.orig_bin_ids = []
        hic.orig_cut_intervals = []
        hic._masked_ids = np.array([], dtype=np.int64)
        hic.correction_factors = None
        log.warning("*WARNING*: The matrix is probably a corrected matrix that contains NaN bins. This bins "
                    "can not be merged and are removed. It is preferable to first merge bins in a uncorrected  "
                    "matrix and then correct the matrix. Correction factors, if present, are removed as well.")
    return hic


def merge_groups(cut_intervals, num_bins):
    """which consecutive bins are merged and the intervals of the merged bins (hicMergeMatrixBins.py:228-262):
    groups of ``num_bins`` inside a chromosome; a chromosome's trailing group with fewer than num_bins/2 bins is
    dropped (bin map -1), except the last group of the matrix."""
    chrom, start, end, cov = zip(*cut_intervals)
    n = len(chrom)
    bin_map = np.full(n, -1, dtype=np.int32)
    new_bins = []
    first, count, prev = 0, 0, chrom[0]
    for idx in range(n):
        if (count > 0 and count % num_bins == 0) or chrom[idx] != prev:
            if count < num_bins / 2:
                log.debug("{} has few bins ({}). Skipping it\n".format(prev, count))
            else:
                bin_map[first:idx] = len(new_bins)
                new_bins.append((chrom[first], start[first], end[idx - 1], float(np.mean(cov[first:idx]))))
            first, count = idx, 0
        prev = chrom[idx]
        count += 1
    bin_map[first:n] = len(new_bins)
    new_bins.append((chrom[n - 1], start[first], end[n - 1], float(np.mean(cov[first:]))))
    return bin_map, new_bins


def merge_bins(hic, num_bins, device=0):
    """Merge ``num_bins`` consecutive bins of the hiCMatrix ``hic`` in place and return it."""
    if num_bins > 1024:
        raise ValueError("--numBins is limited to 1024")
    hic = remove_nans_if_needed(hic)
    bin_map, new_bins = merge_groups(hic.cut_intervals, num_bins)
    dtype = hic.matrix.dtype
    with DeviceCSR.from_scipy(hic.matrix, device=device) as dev:
        dev.merge_bins(bin_map, len(new_bins))
        merged = dev.download()
    if dtype.kind in "iu":
        merged = merged.astype(dtype)        # the reference keeps the input dtype (reduceMatrix.py:213)
    hic.matrix = merged
    hic.cut_intervals = new_bins
    hic._index_intervals()
    hic.nan_bins = np.flatnonzero(np.asarray(hic.matrix.sum(axis=0)).ravel() == 0)
    return hic


def running_window_merge(hic, num_bins, device=0):
    """'running window' merge that keeps the resolution (hicMergeMatrixBins.py:88-186): with U = upper triangle,
    new = sum over (dr, dc) in [-h, h]^2 of U shifted by (dr, dc); result = triu(new) mirrored.  ``num_bins`` must be odd."""
    hic = remove_nans_if_needed(hic)
    if num_bins == 1:
        return hic
    assert num_bins % 2 == 1, "num_bins has to be an odd number"
    half = (num_bins - 1) // 2
    dtype = hic.matrix.dtype
    with DeviceCSR.from_scipy(hic.matrix, device=device) as base:
        base.keep_upper()
        total = base.shifted_copy(0, 0)
        try:
            for dr in range(-half, half + 1):
                for dc in range(-half, half + 1):
                    if dr == 0 and dc == 0:
                        continue
                    with base.shifted_copy(dr, dc) as term:
                        total.add(term)
            total.keep_upper()
            total.symmetrize_upper()
            merged = total.download()
        finally:
            total.close()
    if dtype.kind in "iu":
        merged = merged.astype(dtype)
    hic.matrix = merged
    hic.nan_bins = np.flatnonzero(np.asarray(hic.matrix.sum(axis=0)).ravel() == 0)
    return hic


def main(args=None):
    args = parse_arguments().parse_args(args)
    hic = hm.hiCMatrix(args.matrix)
    if args.runningWindow:
        merged = running_window_merge(hic, args.numBins, device=args.device)
    else:
        merged = merge_bins(hic, args.numBins, device=args.device)
    nan_bins = merged.nan_bins
    merged.save(args.outFileName)
    return nan_bins


if __name__ == "__main__":
    main()
