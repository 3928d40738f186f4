"""ctypes loader for the plain-C CPU restatement (oracle/hic_cpu.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhic_cpu.so")


class SynthParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("nchrom", ctypes.c_int32), ("band", ctypes.c_int32),
                ("n_trans", ctypes.c_int32), ("reserved", ctypes.c_int32), ("d0", ctypes.c_double),
                ("gamma", ctypes.c_double), ("lambda0", ctypes.c_double), ("low_frac", ctypes.c_double),
                ("empty_frac", ctypes.c_double), ("lambda_trans", ctypes.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = ctypes.CDLL(LIB)
        vp = ctypes.c_void_p
        L.ice_loop_coo.restype = ctypes.c_int
        L.ice_loop_coo.argtypes = [ctypes.c_int64, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int, ctypes.c_double,
                                   ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.ice_loop_coo_mt.restype = ctypes.c_int
        L.ice_loop_coo_mt.argtypes = L.ice_loop_coo.argtypes
        L.kr_matvec.restype = None
        L.kr_matvec.argtypes = [ctypes.c_int64, vp, vp, vp, vp, vp, vp, vp, ctypes.c_double, vp, ctypes.c_int]
        L.ice_loop_csr_masked_mt.restype = ctypes.c_int
        L.ice_loop_csr_masked_mt.argtypes = [ctypes.c_int64, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_double)]
        L.row_stats_csr_mt.restype = None
        L.row_stats_csr_mt.argtypes = [ctypes.c_int64, vp, vp, vp, vp, vp, vp, ctypes.c_int]
        L.kr_matvec_mt.restype = None
        L.kr_matvec_mt.argtypes = [ctypes.c_int64, vp, vp, vp, vp, vp, vp, vp, ctypes.c_double, vp, ctypes.c_int]
        L.np_sum_f64.restype = ctypes.c_double
        L.np_sum_f64.argtypes = [vp, ctypes.c_int64]
        L.synth_generate.restype = ctypes.c_int64
        L.synth_generate.argtypes = [ctypes.POINTER(SynthParams), vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def synth_matrix(chrom_bins, seed=0, band=2000, d0=200.0, gamma=1.0, lambda0=40.0, n_trans=8, low_frac=0.02,
                 empty_frac=0.005, lambda_trans=2.0):
    """Same matrix as DeviceCSR.synthetic(...) / oracle.synth_ref.synth_matrix(...), generated in C."""
    cb = np.ascontiguousarray(chrom_bins, dtype=np.int64)
    n = int(cb.sum())
    p = SynthParams(seed=seed, nchrom=len(cb), band=band, n_trans=n_trans, reserved=0, d0=d0, gamma=gamma,
                    lambda0=lambda0, low_frac=low_frac, empty_frac=empty_frac, lambda_trans=lambda_trans)
    indptr = np.zeros(n + 1, dtype=np.int64)
    nnz = lib().synth_generate(ctypes.byref(p), _p(cb), _p(indptr), None, None)
    idx = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    lib().synth_generate(ctypes.byref(p), _p(cb), _p(indptr), _p(idx), _p(val))
    return sparse.csr_matrix((val, idx, indptr), shape=(n, n))


def ice_loop(matrix, max_iter, tol=0.0, threads=1):
    """The reference's hot loop in C on COO triplets.  Returns (bias, iterations, deviation, scaled data)."""
    coo = sparse.coo_matrix(matrix)
    row = np.ascontiguousarray(coo.row, dtype=np.int32)
    col = np.ascontiguousarray(coo.col, dtype=np.int32)
    data = np.ascontiguousarray(coo.data, dtype=np.float64).copy()
    bias = np.ones(matrix.shape[0], dtype=np.float64)
    dev = ctypes.c_double()
    it = lib().ice_loop_coo(matrix.shape[0], len(data), _p(row), _p(col), _p(data), _p(bias), int(max_iter),
                            float(tol), int(threads), ctypes.byref(dev))
    return bias, it, dev.value, data


def ice_loop_mt(matrix, max_iter, tol=0.0, threads=0):
    """ice_loop on all host cores (pthreads; row sums in the serial order, so the results equal ice_loop's bit for bit).
    Returns (bias, iterations, deviation, scaled data, seconds spent inside the C loop)."""
    import os
    import time
    coo = sparse.coo_matrix(sparse.csr_matrix(matrix))          # row-sorted triplets
    row = np.ascontiguousarray(coo.row, dtype=np.int32)
    col = np.ascontiguousarray(coo.col, dtype=np.int32)
    data = np.ascontiguousarray(coo.data, dtype=np.float64).copy()
    bias = np.ones(matrix.shape[0], dtype=np.float64)
    dev = ctypes.c_double()
    threads = int(threads) or (os.cpu_count() or 1)
    t0 = time.time()
    it = lib().ice_loop_coo_mt(matrix.shape[0], len(data), _p(row), _p(col), _p(data), _p(bias), int(max_iter), float(tol),
                               threads, ctypes.byref(dev))
    return bias, it, dev.value, data, time.time() - t0


def kr_matvec(csr, u, x, v=None, p=None, eps=1e-5, threads=0):
    n = csr.shape[0]
    out = np.empty(n, dtype=np.float64)
    ip = np.ascontiguousarray(csr.indptr, dtype=np.int64)
    ix = np.ascontiguousarray(csr.indices, dtype=np.int32)
    lib().kr_matvec(n, _p(ip), _p(ix), _p(np.ascontiguousarray(csr.data, dtype=np.float64)), _p(u), _p(x), _p(v),
                    _p(p), float(eps), _p(out), int(threads))
    return out


def _csr_arrays(csr):
    """(indptr int64, indices int32, data float64) views of a canonical scipy CSR (copies only where the dtype differs)"""
    return (np.ascontiguousarray(csr.indptr, dtype=np.int64), np.ascontiguousarray(csr.indices, dtype=np.int32),
            np.ascontiguousarray(csr.data, dtype=np.float64))


def row_stats_masked(csr, keep=None, threads=0):
    """(row sums, diagonal) of csr[keep][:, keep] in the ORIGINAL frame (masked bins -> 0), all host cores."""
    import os
    ip, ix, da = _csr_arrays(csr)
    n = csr.shape[0]
    rs = np.zeros(n)
    dg = np.zeros(n)
    k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    lib().row_stats_csr_mt(n, _p(ip), _p(ix), _p(da), _p(k), _p(rs), _p(dg), int(threads) or (os.cpu_count() or 1))
    return rs, dg


def ice_loop_masked(csr, keep, max_iter, tol=1e-5, threads=0, inplace=False):
    """iterativeCorrection.py:40-74 on csr[keep][:, keep] without building the sub-matrix (see ice_loop_csr_masked_mt).
    Returns (bias of the kept bins, iterations, deviation, scaled data array in the ORIGINAL pattern, seconds)."""
    import os
    import time
    ip, ix, da = _csr_arrays(csr)
    if not inplace and da is csr.data:
        da = da.copy()
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    n = csr.shape[0]
    bias = np.ones(n)
    dev = ctypes.c_double()
    t0 = time.time()
    it = lib().ice_loop_csr_masked_mt(n, _p(ip), _p(ix), _p(da), _p(keep), _p(bias), int(max_iter), float(tol),
                                      int(threads) or (os.cpu_count() or 1), ctypes.byref(dev))
    return bias[keep.astype(bool)], it, dev.value, da, time.time() - t0


def kr_matvec_mt(csr_arrays, n, u, x, v=None, p=None, eps=1e-5, threads=0):
    import os
    ip, ix, da = csr_arrays
    out = np.empty(n, dtype=np.float64)
    lib().kr_matvec_mt(n, _p(ip), _p(ix), _p(da), _p(u), _p(x), _p(v), _p(p), float(eps), _p(out),
                       int(threads) or (os.cpu_count() or 1))
    return out
