"""Drop-in for the third-party ``krbalancing`` module the reference imports with
``from krbalancing import *`` (hicexplorer/hicCorrectMatrix.py:17) and uses at
hicCorrectMatrix.py:718-732 (``--perchr``) and :744-757 (genome-wide):

    kr = kr_balancing(rows, cols, nnz, indptr_int64, indices_int64, data_float64)
    kr.computeKR()
    x = kr.get_normalisation_vector(True).todense()      # (n, 1) np.matrix
    m = kr.get_normalised_matrix(True)                   # sparse, upper triangular

Same object protocol and statefulness: the first getter called with
``rescale=True`` rescales ``x`` once (so the balanced matrix keeps the original
grand total) and prints "normalisation factor is ..." like the C++ module; later
``rescale=False`` calls return the already-rescaled vector, which is what the
reference's ``--perchr`` + ``.h5`` branch relies on (:727-732).

One deliberate difference: ``get_normalised_matrix`` is idempotent here (the
C++ object rescales its stored matrix in place, so a second call would scale
twice; the reference never calls it twice).

Runs on the GPU through libhicbalance.so; no CPU fallback.
"""
import numpy as np
from scipy import sparse

from . import _lib
from .store import KR_DELTA, KR_DELTA_UPPER, KR_DIAG_EPS, KR_TOL, DeviceCSR

__all__ = ["kr_balancing"]


class kr_balancing(object):
    def __init__(self, rows, cols, nnz, indptr, indices, data, device=0):
        if rows != cols:
            raise ValueError("kr_balancing needs a square matrix")
        _lib.require_device()
        self._n = int(rows)
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices)
        data = np.ascontiguousarray(data, dtype=np.float64)
        m = sparse.csr_matrix((data, indices, indptr), shape=(rows, cols))
        self._dev = DeviceCSR.from_scipy(m, device=device)
        self._dev.pack_values()    # lossless narrow copy of integer counts for the mat-vec kernel; results unchanged
        self._x = None
        self.rescaled = False
        self.outer_iterations = 0
        self.matvecs = 0
        self.residual = np.nan

    def computeKR(self):
        self._x, self.outer_iterations, self.matvecs, self.residual = self._dev.kr(
            tol=KR_TOL, delta=KR_DELTA, Delta=KR_DELTA_UPPER, diag_eps=KR_DIAG_EPS)

    def _need_x(self):
        if self._x is None:
            raise RuntimeError("call computeKR() first")

    def _rescale_once(self, rescale):
        if rescale and not self.rescaled:
            self._x, factor = self._dev.kr_rescale(self._x, KR_DIAG_EPS)
            print("normalisation factor is {}".format(factor))
            self.rescaled = True

    def get_normalisation_vector(self, rescale):
        """sparse (n x 1) column holding x (callers do ``.todense()``)."""
        self._need_x()
        self._rescale_once(rescale)
        return sparse.csc_matrix(self._x.reshape(-1, 1))

    def get_normalised_matrix(self, rescale):
        """upper triangle (diagonal always stored) of (A + 1e-5 I) o (x x^T)."""
        self._need_x()
        self._rescale_once(rescale)
        return self._dev.kr_scaled_upper(self._x, KR_DIAG_EPS)      # CSR: what every caller converts to anyway

    def close(self):
        if self._dev is not None:
            self._dev.close()
            self._dev = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
