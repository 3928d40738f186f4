"""ctypes binding of libhicbalance.so (include/hicbalance.h).

There is NO CPU fallback: if the shared library is missing the import of any
product entry point raises, and every compute call raises ``NoDeviceError`` when
no CUDA device is visible.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HICBALANCE_LIB") or os.path.join(_PKG, "libhicbalance.so")  # env: tuning variants

HB_OK = 0
HB_ERR_NOT_SYMMETRIC = 1
HB_ERR_OVERFLOW_ITER = 2
HB_ERR_OVERFLOW_FINAL = 3
HB_ERR_KR_BREAKDOWN = 4
HB_ERR_INVALID_ARG = 5
HB_ERR_NO_DEVICE = 6
HB_ERR_CUDA = 100


class HicBalanceError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libhicbalance error %d: %s" % (code, message))
        self.code = code


class NoDeviceError(HicBalanceError):
    pass


class OverflowIter(HicBalanceError):
    """a rescaled value exceeded 1e100 (iterativeCorrection.py:58-62)"""


class OverflowFinal(HicBalanceError):
    """a final value exceeded 1e10 (iterativeCorrection.py:80-84)"""


class SynthParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("nchrom", ctypes.c_int32), ("band", ctypes.c_int32),
                ("n_trans", ctypes.c_int32), ("reserved", ctypes.c_int32), ("d0", ctypes.c_double),
                ("gamma", ctypes.c_double), ("lambda0", ctypes.c_double), ("low_frac", ctypes.c_double),
                ("empty_frac", ctypes.c_double), ("lambda_trans", ctypes.c_double)]


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)

c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_dbl = ctypes.c_double
c_int = ctypes.c_int
vp = ctypes.c_void_p
P = ctypes.POINTER

# name -> (restype, argtypes); must list every symbol include/hicbalance.h declares
SIGNATURES = {
    "hb_abi_version": (c_int, []),
    "hb_last_error": (ctypes.c_char_p, []),
    "hb_device_count": (c_int, []),
    "hb_pinned_alloc": (c_int, [P(vp), c_u64]),
    "hb_pinned_free": (c_int, [vp]),
    "hb_launch_count": (c_u64, []),
    "hb_csr_create": (c_int, [P(vp), c_i64, c_i64, c_i64, c_i64, vp, vp, c_int, vp, c_int]),
    "hb_csr_create_upper": (c_int, [P(vp), c_i64, c_i64, vp, vp, vp, c_int]),
    "hb_csr_create_typed": (c_int, [P(vp), c_i64, c_i64, c_i64, c_i64, vp, vp, c_int, vp, c_int, c_int]),
    "hb_csr_create_upper_typed": (c_int, [P(vp), c_i64, c_i64, vp, vp, vp, c_int, c_int]),
    "hb_dev_csr_wrap": (c_int, [P(vp), c_i64, c_i64, c_i64, c_i64, vp, vp, vp, c_int]),
    "hb_csr_destroy": (c_int, [vp]),
    "hb_csr_shape": (c_int, [vp, P(c_i64), P(c_i64), P(c_i64), P(c_i64)]),
    "hb_csr_download": (c_int, [vp, vp, vp, vp]),
    "hb_dev_csr_pointers": (c_int, [vp, P(vp), P(vp), P(vp)]),
    "hb_sync": (c_int, [vp]),
    "hb_set_comm": (c_int, [vp, c_int, c_int, vp, vp]),
    "hb_peer_export": (c_int, [vp, c_int, c_int, vp]),
    "hb_peer_attach": (c_int, [vp, vp]),
    "hb_peer_attach_inprocess": (c_int, [vp, vp]),
    "hb_peer_active": (c_int, [vp]),
    "hb_set_segments": (c_int, [vp, c_int, vp]),
    "hb_csr_keep_cis": (c_int, [vp]),
    "hb_scrub": (c_int, [vp, vp]),
    "hb_csr_pack_values": (c_int, [vp, P(c_int)]),
    "hb_row_stats": (c_int, [vp, vp, vp]),
    "hb_row_stats_cis": (c_int, [vp, vp, vp]),
    "hb_mad_filter": (c_int, [vp, c_dbl, c_dbl, vp, vp, vp, vp]),
    "hb_csr_compact": (c_int, [vp, vp]),
    "hb_diag_zero": (c_int, [vp]),
    "hb_csr_keep_upper": (c_int, [vp]),
    "hb_csr_keep_rows": (c_int, [vp, c_i64, c_i64]),
    "hb_csr_symmetrize_upper": (c_int, [vp]),
    "hb_csr_shifted_copy": (c_int, [P(vp), vp, c_i64, c_i64]),
    "hb_csr_merge_bins": (c_int, [vp, vp, c_i64]),
    "hb_symmetry": (c_int, [vp, P(c_dbl)]),
    "hb_dev_extract_cols": (c_int, [vp, c_i64, c_i64, P(c_i64), vp, vp, vp]),
    "hb_dev_partner_sums": (c_int, [vp, c_i64, vp, vp, vp, P(c_dbl)]),
    "hb_block_correlation": (c_int, [vp, c_i64, c_i64, c_int, P(c_i64), vp, vp, vp]),
    "hb_csr_ingest_info": (c_int, [vp, P(c_int), P(c_i64)]),
    "hb_ice_run": (c_int, [vp, c_int, c_dbl, vp, vp, vp]),
    "hb_ice_scaled_values": (c_int, [vp, vp, P(c_dbl)]),
    "hb_ice_scaled_upper": (c_int, [vp, vp, c_i64, P(c_i64), vp, vp, vp]),
    "hb_csr_value_stats": (c_int, [vp, c_int, P(c_dbl), P(c_dbl), P(c_dbl)]),
    "hb_csr_normalize": (c_int, [vp, c_int, c_dbl, c_dbl, c_dbl, c_int]),
    "hb_csr_eliminate_zeros": (c_int, [vp]),
    "hb_csr_add": (c_int, [vp, vp]),
    "hb_distance_stats": (c_int, [vp, vp, vp]),
    "hb_obs_exp_apply": (c_int, [vp, vp, c_int, c_int, vp, vp]),
    "hb_trans_order_stats": (c_int, [vp, vp, c_int, c_i64, P(c_i64), P(c_dbl), P(c_dbl)]),
    "hb_trans_clip": (c_int, [vp, vp, c_int, c_dbl, P(c_i64)]),
    "hb_ice_scaled_row_sums": (c_int, [vp, vp]),
    "hb_dev_ice_begin": (c_int, [vp, vp]),
    "hb_dev_ice_loop": (c_int, [vp, c_int, c_dbl, vp]),
    "hb_loop_progress": (c_int, [vp, vp]),
    "hb_dev_ice_spmv": (c_int, [vp, vp, c_int, c_int, vp]),
    "hb_dev_ice_update": (c_int, [vp, vp, c_int, c_dbl, vp]),
    "hb_dev_ice_status": (c_int, [vp, vp, vp]),
    "hb_dev_ice_finish": (c_int, [vp, vp]),
    "hb_dev_ice_emit": (c_int, [vp, vp, vp, vp]),
    "hb_dev_ice_vectors": (c_int, [vp, P(vp), P(vp)]),
    "hb_ice_results": (c_int, [vp, vp, vp, vp]),
    "hb_kr_run": (c_int, [vp, c_dbl, c_dbl, c_dbl, c_dbl, c_int, vp, P(c_i32), P(c_i32), P(c_dbl)]),
    "hb_kr_rescale": (c_int, [vp, c_dbl, vp, P(c_dbl)]),
    "hb_kr_scaled_upper": (c_int, [vp, c_dbl, vp, P(c_i64), vp, vp, vp]),
    "hb_dev_spmv": (c_int, [vp, vp, c_dbl, vp, vp, vp, vp, c_int, vp]),
    "hb_synth_create": (c_int, [P(vp), P(SynthParams), vp, c_i64, c_i64, c_int]),
}

_lib = None


def lib():
    """The loaded library (loads on first use; raises if it was never built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(hicexplorer_b200 has no CPU fallback)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    msg = lib().hb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code):
    if code == HB_OK:
        return
    msg = last_error()
    if code == HB_ERR_NO_DEVICE:
        raise NoDeviceError(code, msg)
    if code == HB_ERR_OVERFLOW_ITER:
        raise OverflowIter(code, msg)
    if code == HB_ERR_OVERFLOW_FINAL:
        raise OverflowFinal(code, msg)
    if code == HB_ERR_INVALID_ARG:
        raise ValueError("libhicbalance: " + msg)
    raise HicBalanceError(code, msg)


def require_device():
    n = lib().hb_device_count()
    if n <= 0:
        raise NoDeviceError(HB_ERR_NO_DEVICE, "no CUDA device visible; hicexplorer_b200 has no CPU fallback")
    return n


def ptr(arr):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(vp)


def as_c(arr, dtype):
    return np.ascontiguousarray(arr, dtype=dtype)
