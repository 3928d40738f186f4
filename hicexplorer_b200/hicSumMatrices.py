"""``hicSumMatrices`` on a B200: same flags, checks and output as the reference tool
(hicexplorer/hicSumMatrices.py:9-74); the sparse additions run on the GPU (``hb_csr_add``: union of the two sorted
rows, values added, exact zeros dropped like scipy's ``csr + csr``).  SURVEY.md 8f, rank 3."""
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
        add_help=False, description='Adds Hi-C matrices of the same size. Format has to be hdf5 or npz')
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrices', '-m', help='Space-delimited names of the matrices to add. The matrices must have the '
                     'same shape/size.', metavar='matrix.h5', nargs='+', required=True)
    req.add_argument('--outFileName', '-o', help='File name to save the resulting matrix. The output is also a .h5 '
                     'file.', required=True)
    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--help', '-h', action='help', help='show this help message and exit')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    opt.add_argument('--device', help='CUDA device index.', type=int, default=0)
    return parser


def main(args=None):
    args = parse_arguments().parse_args(args)
    hic = hm.hiCMatrix(args.matrices[0])
    nan_bins = set(int(b) for b in hic.nan_bins)
    dtype = hic.matrix.dtype
    with DeviceCSR.from_scipy(hic.matrix, device=args.device) as total:
        for path in args.matrices[1:]:
            other = hm.hiCMatrix(path)
            if hic.chrBinBoundaries != other.chrBinBoundaries:
                log.error("The two matrices have different chromosome order. Use the tool `hicConvertFormat` to change "
                          "the order.\n{}: {}\n{}: {}".format(args.matrices[0], list(hic.chrBinBoundaries), path,
                                                              list(other.chrBinBoundaries)))
                sys.exit(1)
            try:
                if other.matrix.shape != hic.matrix.shape:
                    raise ValueError("inconsistent shapes")
                with DeviceCSR.from_scipy(other.matrix, device=args.device) as addend:
                    total.add(addend)
                dtype = np.result_type(dtype, other.matrix.dtype)
                if len(other.nan_bins):
                    nan_bins = nan_bins.union(int(b) for b in other.nan_bins)
            except Exception:
                log.exception("\nMatrix {} seems to be corrupted or of different shape".format(path))
                sys.exit(1)
        summed = total.download()
    if dtype.kind in "iu":
        summed = summed.astype(dtype)
    # save only the upper triangle of the symmetric matrix
    hic.setMatrixValues(summed)
    hic.maskBins(sorted(nan_bins))
    hic.save(args.outFileName)


if __name__ == "__main__":
    main()
