"""numpy restatement of Knight-Ruiz balancing as the reference uses it.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The arithmetic is NOT under ``/root/reference``: the reference calls the
third-party C++/Eigen/pybind11 package ``krbalancing`` (pinned only as
``krbalancing >= 0.0.5``, ``requirements.txt:20`` / ``setup.py:113``; upstream
project deeptools/Knight-Ruiz-Matrix-balancing-algorithm) at
``hicexplorer/hicCorrectMatrix.py:17, 718-732, 744-757``.  This file restates
its published algorithm -- Knight & Ruiz (2012) "A fast algorithm for matrix
balancing", routine ``bnewt`` (inexact Newton with a CG inner solve and a
cone-boundary step limiter) -- with the package's constants and its two
conventions that shape the output:

* the constructor balances ``A + 1e-5*I`` (every bin gets a stored diagonal;
  empty bins balance to ``x = 1/sqrt(1e-5)``),
* ``rescale=True`` divides ``x`` by ``sqrt(sum(A o x x^T) / sum(A))`` so that
  the balanced matrix keeps the original grand total.

PARITY PINNING.  The reference's own tests hold one known answer for this
path: ``test_data/hicCorrectMatrix/small_test_matrix_KRcorrected_chrUextra_chr3LHet.h5``
(``test/general/test_hicCorrectMatrix.py:52-69``, compared to 5 decimals).
``tests/test_oracle.py`` checks this restatement against it (same sparsity
pattern, max abs error < 1e-7) and against the ``bins/weight`` column of
``kr_partial.cool`` (uniform ratio).  Beyond those 5 decimals KR parity is
unpinned by the reference (no krbalancing build exists here: it needs Eigen).
"""
import numpy as np
from scipy import sparse

TOL = 1e-6
DELTA_LO = 0.1
DELTA_HI = 3.0
G = 0.9
ETA_MAX = 0.1
DIAG_EPS = 1e-5


def _augment(csr):
    n = csr.shape[0]
    a = sparse.csr_matrix(csr, dtype=np.float64)
    return sparse.csr_matrix(a + DIAG_EPS * sparse.identity(n, dtype=np.float64, format="csr"))


class _ThreadedOperator(object):
    """(A + 1e-5 I) @ u through oracle/hic_cpu.c's row-parallel mat-vec (all host cores) -- for inputs of 1e8+ entries,
    where scipy's single-threaded product and the augmented copy of the matrix cost minutes and gigabytes.  The diagonal
    term is added per row (acc + eps * u_i) instead of being stored, a difference of one rounding per row."""

    def __init__(self, csr):
        from . import hic_cpu
        self._hc = hic_cpu
        self._arrays = hic_cpu._csr_arrays(csr)
        self.shape = csr.shape
        self._ones = np.ones(csr.shape[0])

    def __matmul__(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        return self._hc.kr_matvec_mt(self._arrays, self.shape[0], u, self._ones, eps=DIAG_EPS)


def kr_balance(csr, tol=TOL, max_outer=100000, threaded=False):
    """Returns dict(x, outer, matvecs, residual, A) for A = csr + 1e-5*I."""
    a = _ThreadedOperator(csr) if threaded else _augment(csr)
    n = a.shape[0]
    x = np.ones(n)
    v = x * (a @ x)
    rk = 1.0 - v
    rho_km1 = float(rk @ rk)
    rho_km2 = rho_km1
    rout = rold = rho_km1
    eta = ETA_MAX
    stop_tol = tol * 0.5
    rt = tol * tol
    outer = 0
    matvecs = 0
    boundary_steps = 0
    z = p = None
    while rout > rt and outer < max_outer:
        outer += 1
        k = 0
        y = np.ones(n)
        innertol = max(eta * eta * rout, rt)
        while rho_km1 > innertol:
            k += 1
            if k == 1:
                z = rk / v
                p = z.copy()
                rho_km1 = float(rk @ z)
            else:
                beta = rho_km1 / rho_km2
                p = z + beta * p
            w = x * (a @ (x * p)) + v * p
            alpha = rho_km1 / float(p @ w)
            ap = alpha * p
            ynew = y + ap
            if ynew.min() <= DELTA_LO:
                if DELTA_LO == 0:
                    break
                sel = ap < 0
                gamma = np.min((DELTA_LO - y[sel]) / ap[sel])
                y = y + gamma * ap
                boundary_steps += 1
                break
            if ynew.max() >= DELTA_HI:
                sel = ynew > DELTA_HI
                gamma = np.min((DELTA_HI - y[sel]) / ap[sel])
                y = y + gamma * ap
                boundary_steps += 1
                break
            y = ynew
            rk = rk - alpha * w
            rho_km2 = rho_km1
            z = rk / v
            rho_km1 = float(rk @ z)
        x = x * y
        v = x * (a @ x)
        rk = 1.0 - v
        rho_km1 = float(rk @ rk)
        rout = rho_km1
        matvecs += k + 1
        rat = rout / rold
        rold = rout
        res_norm = np.sqrt(rout)
        eta_o = eta
        eta = G * rat
        if G * eta_o * eta_o > 0.1:
            eta = max(eta, G * eta_o * eta_o)
        eta = max(min(eta, ETA_MAX), stop_tol / res_norm)
    return dict(x=x, outer=outer, matvecs=matvecs, residual=float(np.sqrt(rout)), A=a, boundary_steps=boundary_steps)


def rescale_factor(a, x):
    """sqrt(sum_ij A_ij x_i x_j / sum_ij A_ij), accumulated over the upper
    triangle with off-diagonals counted twice (krbalancing's convention)."""
    up = sparse.triu(a, 0).tocoo()
    off = up.row != up.col
    orig = up.data[~off].sum() + 2.0 * up.data[off].sum()
    scaled = up.data * x[up.row] * x[up.col]
    norm = scaled[~off].sum() + 2.0 * scaled[off].sum()
    return float(np.sqrt(norm / orig))


def normalised_upper(a, x):
    """Upper triangle (diagonal included) of A o (x x^T), CSR."""
    up = sparse.triu(a, 0).tocoo()
    data = up.data * x[up.row] * x[up.col]
    out = sparse.csr_matrix((data, (up.row, up.col)), shape=a.shape)
    out.sort_indices()
    return out


def kr_pipeline(csr, rescale=True):
    """What hicCorrectMatrix's KR branch obtains from the object
    (hicCorrectMatrix.py:744-757): the (rescaled) vector and the upper-triangular
    normalised matrix."""
    res = kr_balance(csr)
    x = res["x"]
    factor = 1.0
    if rescale:
        factor = rescale_factor(res["A"], x)
        x = x / factor
    res.update(x_out=x, factor=factor, upper=normalised_upper(res["A"], x))
    return res
