"""``hicNormalize`` on a B200: same flags and output as the reference tool (hicexplorer/hicNormalize.py:13-153), the
arithmetic on the GPU (``hb_csr_value_stats`` / ``hb_csr_normalize`` / ``hb_csr_eliminate_zeros``).

The step the reference's documentation places before the correction when several samples are compared
(SURVEY.md 8f, rank 3).  Like the reference the values are computed and written in float32
(``matrix.data.astype(np.float32)``, :76/:105/:131).  The scalar operand (min/max range, sum ratio,
``--multiplicativeValue``) enters the float32 operation as a float32, which is what the numpy the reference was
written against did with ``np.divide(float32_array, float64_scalar)`` and what its golden files contain.
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
        formatter_class=argparse.RawDescriptionHelpFormatter, add_help=False,
        description='Normalizes given matrices either to the smallest given read number of all matrices or to 0 - 1 '
                    'range. However, it does NOT compute the contact probability.')
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrices', '-m', help='The matrix (or multiple matrices) to get information about. '
                     'HiCExplorer supports the following file formats: h5 (native HiCExplorer format) and cool.',
                     nargs='+', required=True)
    req.add_argument('--normalize', '-n', help='Normalize to a) 0 to 1 range, b) all matrices to the lowest read '
                     'count of the given matrices (Default: %(default)s).',
                     choices=['norm_range', 'smallest', 'multiplicative'], default='smallest', required=True)
    req.add_argument('--outFileName', '-o', help='Output file name for the Hi-C matrix.', metavar='FILENAME', nargs='+',
                     required=True)
    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--multiplicativeValue', '-mv', type=float, help='Value to multiply if --normalize is set to '
                     'multiplicative. (Default: %(default)s).', default=1)
    opt.add_argument('--setToZeroThreshold', '-sz', default=0.0, type=float,
                     help='A threshold to set all values after normalization to 0 if smaller this threshold. Default '
                          'value is 0 i.e. there is no effect.')
    opt.add_argument('--help', '-h', action='help', help='show this help message and exit')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    opt.add_argument('--device', help='CUDA device index.', type=int, default=0)
    return parser


def _store_result(hic, dev):
    dev.eliminate_zeros()                              # hicNormalize.py:93 / :97
    m = dev.download()
    m.data = m.data.astype(np.float32)                 # exact: the values were produced in float32
    hic.matrix = m
    hic.value_dtype = np.float32


def normalize_matrices(hics, mode, multiplicative_value=1.0, threshold=0.0, device=0):
    """the body of hicNormalize.main (:75-153) for already loaded matrices; modifies and returns them"""
    if mode == 'smallest':
        sums = []
        devs = []
        for hic in hics:
            dev = DeviceCSR.from_scipy(hic.matrix, device=device)
            devs.append(dev)
            sums.append(dev.value_stats(scrub=False)[0])           # `hic_ma.matrix.sum()` of the raw values (:70)
        argmin = int(np.argmin(sums))
        for i, (hic, dev) in enumerate(zip(hics, devs)):
            if i != argmin:
                dev.normalize('divide', a=np.float32(sums[i] / sums[argmin]), threshold=threshold, scrub_first=True)
            else:
                dev.normalize('cast', threshold=threshold, scrub_first=False)
            _store_result(hic, dev)
            dev.close()
        return hics
    for hic in hics:
        with DeviceCSR.from_scipy(hic.matrix, device=device) as dev:
            if mode == 'norm_range':
                _s, mn, mx = dev.value_stats(scrub=True)
                diff = np.float32(mx) - np.float32(mn)             # float32 subtraction, like max_value - min_value (:84)
                dev.normalize('norm_range', a=mn, b=diff, threshold=threshold, scrub_first=True)
            else:
                dev.normalize('multiply', a=np.float32(multiplicative_value), threshold=threshold, scrub_first=True)
            _store_result(hic, dev)
    return hics


def main(args=None):
    args = parse_arguments().parse_args(args)
    hics = [hm.hiCMatrix(m) for m in args.matrices]
    normalize_matrices(hics, args.normalize, args.multiplicativeValue, args.setToZeroThreshold, device=args.device)
    for hic, out in zip(hics, args.outFileName):
        hic.save(out, pApplyCorrection=False)


if __name__ == "__main__":
    main()
