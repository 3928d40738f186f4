"""``hicSumMatrices`` on a B200.

Drop-in for the reference tool of the same name (hicexplorer/hicSumMatrices.py:9-74): identical flags
(``--matrices/-m``, ``--outFileName/-o``), identical refusal of inputs whose chromosome layout differs, identical
output (the sum with the union of the inputs' NaN bins masked).  The additions themselves run on the GPU through
``hb_csr_add`` -- each row of the running total is merged with the matching row of the next matrix, equal columns
are added, and cells that cancel to exactly zero disappear, which is what scipy's ``csr + csr`` does in the reference
(:59).  SURVEY.md 8f, rank 3.
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
    p = argparse.ArgumentParser(add_help=False, description="Element-wise sum of Hi-C matrices that share one bin layout.")
    need = p.add_argument_group("Required arguments")
    need.add_argument("--matrices", "-m", nargs="+", required=True, metavar="matrix.h5",
                      help="Two or more matrices (.h5 or .cool) with the same bins, separated by spaces.")
    need.add_argument("--outFileName", "-o", required=True, help="Where to write the summed matrix.")
    extra = p.add_argument_group("Optional arguments")
    extra.add_argument("--help", "-h", action="help", help="Print this text and leave.")
    extra.add_argument("--version", action="version", version="%(prog)s {}".format(__version__))
    extra.add_argument("--device", type=int, default=0, help="CUDA device to run on.")
    return p


def _same_layout_or_exit(first, first_name, other, other_name):
    if first.chrBinBoundaries == other.chrBinBoundaries:
        return
    log.error("The two matrices have different chromosome order. Use the tool `hicConvertFormat` to change the order.\n"
              "{}: {}\n{}: {}".format(first_name, list(first.chrBinBoundaries), other_name, list(other.chrBinBoundaries)))
    sys.exit(1)


def sum_matrices(paths, device=0):
    """Load ``paths`` one after the other and accumulate them in one device store.  Returns the first matrix object
    with the total installed and the union of all NaN bins masked."""
    base = hm.hiCMatrix(paths[0])
    masked = {int(b) for b in base.nan_bins}
    value_type = base.matrix.dtype
    with DeviceCSR.from_scipy(base.matrix, device=device) as running:
        for path in paths[1:]:
            nxt = hm.hiCMatrix(path)
            _same_layout_or_exit(base, paths[0], nxt, path)
            try:
                if nxt.matrix.shape != base.matrix.shape:
                    raise ValueError("shape %r differs from %r" % (nxt.matrix.shape, base.matrix.shape))
                with DeviceCSR.from_scipy(nxt.matrix, device=device) as term:
                    running.add(term)
            except Exception:
                log.exception("\nMatrix {} seems to be corrupted or of different shape".format(path))
                sys.exit(1)
            value_type = np.result_type(value_type, nxt.matrix.dtype)
            masked.update(int(b) for b in nxt.nan_bins)
        total = running.download()
    base.setMatrixValues(total.astype(value_type) if value_type.kind in "iu" else total)
    base.maskBins(sorted(masked))
    return base


def main(args=None):
    opts = parse_arguments().parse_args(args)
    sum_matrices(opts.matrices, device=opts.device).save(opts.outFileName)


if __name__ == "__main__":
    main()
