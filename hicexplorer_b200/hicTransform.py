"""``hicTransform`` (observed / expected methods) on a B200: same flags and output as the reference tool
(hicexplorer/hicTransform.py:13-260) for ``--method obs_exp | obs_exp_non_zero | obs_exp_lieberman`` with
``--perChromosome``, ``--ligation_factor`` and ``--chromosomes``; SURVEY.md 8f, rank 4.

The two O(nnz) steps run on the GPU: the per-distance sums / counts (``hb_distance_stats``; an
O(n_distances x nnz) mask loop in the reference, utilities.py:424-426) and the division of every entry by its expected
value (``hb_obs_exp_apply``).  The O(n) step in between uses the reference's own numpy expressions, quirks included:
``obs_exp`` divides the distance sums by ``n + 1 - d`` and, like ``obs_exp_lieberman``, looks the expected value up at
``ceil(d / 2)`` (utilities.py:364, 489, 581); the quotient is cast back to the input dtype (integer count matrices are
truncated), ``--perChromosome`` assembles into a float64 ``lil_matrix`` and ``obs_exp_non_zero`` yields float32.

``pearson`` and ``covariance`` (``np.corrcoef`` / ``np.cov`` of the dense matrix or, with ``--perChromosome``, of every
chromosome block; hicTransform.py:105-113, 213-248) densify the block on the device, take the product with one fp64
cuBLAS DGEMM and return the non-zero cells (``hb_block_correlation``).
"""
import argparse
import logging
import sys

import numpy as np

from . import hicmatrix_lite as hm
from ._version import __version__
from .store import DeviceCSR

log = logging.getLogger(__name__)


def parse_arguments(args=None):
    parser = argparse.ArgumentParser(
        formatter_class=argparse.RawDescriptionHelpFormatter, add_help=False,
        description='Converts the (interaction) matrix to an observed/expected matrix or a pearson_correlation matrix.')
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrix', '-m', help='input file. The computation is done per chromosome.', required=True)
    req.add_argument('--outFileName', '-o', help='File name to save the exported matrix.', required=True)
    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--method', '-me', help='Transformation methods to use for input matrix (Default: %(default)s).',
                     choices=['obs_exp', 'obs_exp_lieberman', 'obs_exp_non_zero', 'pearson', 'covariance'],
                     default='obs_exp')
    opt.add_argument('--ligation_factor', help='Multiplies a scaling factor to each entry of the expected matrix to take '
                     'care of the proximity ligation (Homer). Only with obs_exp_non_zero.', action='store_true')
    opt.add_argument('--chromosomes', help='List of chromosomes to be included in the computation.', default=None, nargs='+')
    opt.add_argument('--perChromosome', '-pc', help='Each chromosome is processed individually, inter-chromosomal '
                     'interactions are ignored. Option not valid for obs_exp_lieberman.', action='store_true')
    opt.add_argument('--help', '-h', action='help', help='Show this help message and exit.')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    opt.add_argument('--device', help='CUDA device index.', type=int, default=0)
    return parser


def _scrub(x):
    x[np.isnan(x)] = 0
    x[np.isinf(x)] = 0
    return x


def expected_values(method, sums, counts, offsets):
    """the O(n) step: distance sums -> expected values, per block [offsets[s], offsets[s+1]) (utilities.py:293-338, 364, 432)"""
    expected = np.zeros(len(sums), dtype=np.float64)
    total_len, n_chr = int(offsets[-1] - offsets[0]), len(offsets) - 1
    with np.errstate(divide="ignore", invalid="ignore"):
        for lo, hi in zip(offsets[:-1], offsets[1:]):
            n = int(hi - lo)
            if method == 'obs_exp':
                expected[lo:hi] = _scrub(np.divide(sums[lo:hi], np.arange(n + 1, 1, -1)))
            elif method == 'obs_exp_non_zero':
                expected[lo:hi] = _scrub(np.divide(sums[lo:hi], counts[lo:hi].astype(np.float64)))
            else:
                count_times_i = np.arange(float(n))
                count_times_i *= np.int32(n_chr)
                count_times_i -= np.int32(total_len)
                count_times_i *= np.int32(-1)
                expected[lo:hi] = np.divide(sums[lo:hi], count_times_i)
    return expected


def obs_exp_transform(matrix, method, chr_offsets, per_chromosome=False, ligation_factor=False, device=0):
    """the obs/exp branches of hicTransform.main (:162-215) for a loaded full symmetric matrix; returns the new CSR"""
    in_dtype = matrix.dtype
    blocks = per_chromosome or method == 'obs_exp_lieberman'
    n = matrix.shape[0]
    offsets = np.asarray(chr_offsets, dtype=np.int64) if blocks else np.asarray([0, n], dtype=np.int64)
    with DeviceCSR.from_scipy(matrix, device=device) as dev:
        if blocks:
            dev.set_segments(offsets)
            dev.keep_cis()
        sums, counts = dev.distance_stats()
        expected = expected_values(method, sums, counts, offsets)
        row_sum = seg_total = None
        if method == 'obs_exp_non_zero' and ligation_factor:
            row_sum, _ = dev.row_stats()
            seg_total = np.asarray([row_sum[lo:hi].sum() for lo, hi in zip(offsets[:-1], offsets[1:])])
        if method == 'obs_exp_non_zero':
            out_mode = 2
        else:
            out_mode = 1 if in_dtype.kind in "iu" else (2 if in_dtype == np.float32 else 0)
        dev.obs_exp_apply(expected, 0 if method == 'obs_exp_non_zero' else 1, out_mode, row_sum, seg_total)
        dev.eliminate_zeros()
        out = dev.download()
    if blocks:
        return out                                     # assembled through a float64 lil_matrix in the reference (:161)
    out.data = out.data.astype(np.float32 if method == 'obs_exp_non_zero' else in_dtype)
    return out


def correlation_transform(matrix, method, chr_offsets, per_chromosome=False, device=0):
    """the pearson / covariance branches of hicTransform.main (:213-248) for a loaded full symmetric matrix"""
    import scipy.sparse as sp
    n = matrix.shape[0]
    offsets = np.asarray(chr_offsets, dtype=np.int64) if per_chromosome else np.asarray([0, n], dtype=np.int64)
    with DeviceCSR.from_scipy(matrix, device=device) as dev:
        blocks = [dev.block_correlation(lo, hi, pearson=(method == 'pearson')) for lo, hi in zip(offsets[:-1], offsets[1:])]
    return sp.vstack(blocks, format='csr') if blocks else sp.csr_matrix((n, n), dtype=np.float64)


def main(args=None):
    args = parse_arguments().parse_args(args)
    if not args.outFileName.endswith('.h5') and not args.outFileName.endswith('.cool'):
        log.error('Output filetype not known.')
        log.error('It is: {}'.format(args.outFileName))
        log.error('Accepted is .h5 or .cool')
        sys.exit(1)
    if args.matrix.endswith('cool') and args.chromosomes is not None and len(args.chromosomes) == 1:
        hic_ma = hm.hiCMatrix(pMatrixFile=args.matrix, pChrnameList=args.chromosomes)
    else:
        hic_ma = hm.hiCMatrix(pMatrixFile=args.matrix)
        if args.chromosomes:
            hic_ma.keepOnlyTheseChr([str(c) for c in args.chromosomes])     # file order kept (hicTransform.py:156)
    if args.method in ('pearson', 'covariance'):
        result = correlation_transform(hic_ma.matrix, args.method, hic_ma.chromosome_offsets(), per_chromosome=args.perChromosome,
                                       device=args.device)
    else:
        result = obs_exp_transform(hic_ma.matrix, args.method, hic_ma.chromosome_offsets(), per_chromosome=args.perChromosome,
                                   ligation_factor=args.ligation_factor, device=args.device)
    hic_ma.matrix = result
    hic_ma.value_dtype = result.dtype
    hic_ma.save(args.outFileName, pSymmetric=True, pApplyCorrection=False)


if __name__ == "__main__":
    main()
