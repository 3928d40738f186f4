"""numpy restatement of ``hicNormalize`` and ``hicSumMatrices``.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``normalize()``  <- ``hicexplorer/hicNormalize.py:63-153`` (the three modes, NaN/Inf scrub, threshold, eliminate_zeros)
* ``sum_matrices()`` <- ``hicexplorer/hicSumMatrices.py:48-72``

The reference computes in float32 (``data.astype(np.float32)``) and combines the float32 array with a float64 scalar
(``np.divide(data, adjust_factor)``).  With the numpy it was written against (value-based casting, numpy < 2) that
operation runs in float32; numpy >= 2 would promote it to float64.  The golden files of the reference's tests are
float32 and are reproduced bit for bit by the float32 form, which is what is restated here (explicit ``np.float32``
scalars make the result independent of the installed numpy).

Pinned (tests/test_oracle.py) against the reference's golden files ``test_data/hicNormalize/smallest_{one,two}.h5``,
``norm_range_{one,two}.h5`` and ``small_test_matrix_scaled_by_2.cool``.
"""
import numpy as np
from scipy import sparse


def _scrub32(data):
    data[np.isnan(data)] = 0
    data[np.isinf(data)] = 0
    return data


def _finish(m, threshold):
    _scrub32(m.data)
    m.eliminate_zeros()
    m.data[m.data < np.float32(threshold)] = 0
    m.eliminate_zeros()
    return m


def normalize(matrices, mode, multiplicative_value=1.0, threshold=0.0):
    """``matrices``: list of scipy CSR (full symmetric, as hicmatrix holds them).  Returns new float32 CSR matrices."""
    out = []
    if mode == "smallest":
        sums = [m.sum() for m in matrices]                      # :70, before the float32 cast
        argmin = int(np.argmin(sums))
    for i, m in enumerate(matrices):
        m = sparse.csr_matrix(m, copy=True)
        m.data = m.data.astype(np.float32)
        if mode == "norm_range":                                # :76-97
            _scrub32(m.data)
            mn, mx = np.min(m.data), np.max(m.data)
            diff = np.float32(mx - mn)
            m.data -= mn
            m.data = np.divide(m.data, diff).astype(np.float32)
        elif mode == "smallest":                                # :104-121
            if i != argmin:
                _scrub32(m.data)
                adjust = np.float32(sums[i] / sums[argmin])
                m.data = np.divide(m.data, adjust).astype(np.float32)
        elif mode == "multiplicative":                          # :131-145
            _scrub32(m.data)
            m.data *= np.float32(multiplicative_value)
        else:
            raise ValueError(mode)
        out.append(_finish(m, threshold))
    return out


def sum_matrices(matrices):
    total = matrices[0]
    for m in matrices[1:]:
        total = total + m                                        # :59 (scipy drops cells that sum to 0)
    return sparse.csr_matrix(total)
