"""Drop-in for ``hicexplorer.iterativeCorrection.iterativeCorrection``
(reference: hicexplorer/iterativeCorrection.py:10-86), running on a B200
through libhicbalance.so.  Same signature, same return value
``(corrected CSR float64, bias ndarray)``, same side effects:

* NaNs are zeroed in the CALLER's ``matrix.data`` with the same warning (:28-30),
* ``ValueError("Please provide symmetric matrix!")`` for asymmetric input (:35-36),
* overflow guards end the process with ``SystemExit(1)`` after ``log.error`` (:58-62, :80-84),
* "[iterative correction] N iterations used" is logged with the reference's
  off-by-one (N = loop bodies executed + 1; :41, :73) and only when converged,
* ``verbose`` raises the logger to INFO and reports every 5th pass (:23-24, :63-70).

There is no CPU fallback: without the CUDA library/device this raises.
"""
import logging
import time

import numpy as np
from scipy import sparse

from . import _lib
from ._lib import check, lib, ptr
from .store import DeviceCSR, parallel_copy

log = logging.getLogger(__name__)

_MSG_ITER = ("*Error* matrix correction is producing extremely large values. "
             "This is often caused by bins of low counts. Use a more stringent "
             "filtering of bins.")
_MSG_FINAL = ("*Error* matrix correction produced extremely large values. "
              "This is often caused by bins of low counts. Use a more stringent "
              "filtering of bins.")


def iterativeCorrection(matrix, v=None, M=50, tolerance=1e-5, verbose=False, device=0):
    """ICE (Imakaev et al.) on the GPU.

    :param matrix: a scipy sparse matrix (square, symmetric; any format/dtype)
    :param v: unused (kept for signature compatibility, as in the reference)
    :param M: maximum number of iterations
    :param tolerance: maximum allowed relative deviation of the marginals
    """
    if verbose:
        log.setLevel(logging.INFO)
    _lib.require_device()

    # `if np.isnan(matrix.sum())` (:28): the sum is taken on the device (3 ms for 1.5e9 entries instead of a host pass);
    # only a matrix that does contain NaNs is fixed on the host -- in the CALLER's array, like the reference -- and
    # uploaded again
    dev, m = DeviceCSR.from_scipy_as_uploaded(matrix, device=device)
    if np.isnan(dev.value_stats()[0]):
        dev.close()
        log.warning("[iterative correction] the matrix contains nans, they will be replaced by zeros.")
        data = getattr(matrix, "data", None)
        if isinstance(data, np.ndarray) and data.dtype.kind == "f":
            data[np.isnan(data)] = 0
        if m is not matrix and m.dtype.kind == "f":
            m.data[np.isnan(m.data)] = 0
        dev, m = DeviceCSR.from_scipy_as_uploaded(m, device=device)
    # the result is a matrix of its own, like the reference's (new index arrays): they are copied on a helper thread while the
    # device checks the symmetry, iterates and emits
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1)
    index_copy = pool.submit(lambda: (parallel_copy(m.indices), m.indptr.copy()))
    pool.shutdown(wait=False)
    with dev:
        if dev.symmetry_ratio() > 1e-10:
            index_copy.result()
            raise ValueError("Please provide symmetric matrix!")
        log.info("starting iterative correction")
        dev.pack_values()      # lossless narrow copy of integer counts for the streaming kernel; results unchanged
        try:
            if verbose:
                bias, iters, deviation = _ice_verbose(dev, M, tolerance)
            else:
                bias, iters, deviation = dev.ice(M, tolerance)
        except _lib.OverflowIter:
            log.error(_MSG_ITER)
            raise SystemExit(1)
        if M > 0 and deviation[0] < tolerance:
            log.info("[iterative correction] {} iterations used\n".format(int(iters[0]) + 1))
        try:
            values = dev.ice_scaled_values()
        except _lib.OverflowIter:
            log.error(_MSG_ITER)
            raise SystemExit(1)
        except _lib.OverflowFinal:
            log.error(_MSG_FINAL)
            raise SystemExit(1)
    indices, indptr = index_copy.result()
    corrected = sparse.csr_matrix((values, indices, indptr), shape=m.shape)
    return corrected, bias


def _ice_verbose(dev, M, tolerance):
    """Same loop through the step-level API, reporting every 5th pass like the reference."""
    h = dev.handle
    check(lib().hb_dev_ice_begin(h, None))
    status = np.zeros(4, dtype=np.int32)
    start = time.time()
    dev_now = np.zeros(1)
    for it in range(1, M + 1):
        check(lib().hb_dev_ice_spmv(h, None, 0, 1, None))
        check(lib().hb_dev_ice_update(h, None, 1, float(tolerance), None))
        if it % 5 == 0 or it == M:
            check(lib().hb_dev_ice_status(h, ptr(status), None))
            check(lib().hb_ice_results(h, None, None, ptr(dev_now)))  # synchronises the stream
            if it % 5 == 0:
                est = (float(M - it) * (time.time() - start)) / it
                mnt, sec = divmod(est, 60)
                hrs, mnt = divmod(mnt, 60)
                log.info("pass {} Estimated time {:.0f}:{:.0f}:{:.0f}".format(it, hrs, mnt, sec))
                log.info("max delta - 1 = {} ".format(dev_now[0]))
            if status[0]:
                break
    check(lib().hb_dev_ice_status(h, ptr(status), None))
    dev.sync()
    if status[1]:
        raise _lib.OverflowIter(_lib.HB_ERR_OVERFLOW_ITER, "scaled value exceeded 1e100")
    check(lib().hb_dev_ice_finish(h, None))
    bias = np.empty(dev.n, dtype=np.float64)
    iters = np.zeros(1, dtype=np.int32)
    check(lib().hb_ice_results(h, ptr(bias), ptr(iters), ptr(dev_now)))
    return bias, iters, dev_now
