"""CPU checks of the drop-in boundary: the C-ABI library is built in-tree, loads, exports every symbol
that include/hicbalance.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hicbalance.h")).read()
    return sorted(set(re.findall(r"HB_API\s+[\w\s\*]+?\b(hb_\w+)\s*\(", text)))


def test_library_is_built_and_exports_every_declared_symbol():
    from hicexplorer_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(handle, name), "missing export: " + name
    # the ctypes table binds exactly the declared surface
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().hb_abi_version() == 1


def test_error_string_and_argument_validation_without_gpu():
    from hicexplorer_b200 import _lib
    L = _lib.lib()
    rc = L.hb_set_comm(None, 0, 1, None, None)
    assert rc == _lib.HB_ERR_INVALID_ARG
    assert "null handle" in _lib.last_error()


def test_product_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hicexplorer_b200 import _lib
    from hicexplorer_b200.iterativeCorrection import iterativeCorrection
    from hicexplorer_b200.krbalancing import kr_balancing
    m = sparse.identity(4, format="csr")
    with pytest.raises(_lib.NoDeviceError):
        iterativeCorrection(m)
    with pytest.raises(_lib.NoDeviceError):
        kr_balancing(4, 4, 4, m.indptr.astype(np.int64), m.indices.astype(np.int64), m.data.astype(np.float64))
    h = ctypes.c_void_p()
    ip = np.arange(5, dtype=np.int64)
    rc = _lib.lib().hb_csr_create(ctypes.byref(h), 4, 4, 0, 4, _lib.ptr(ip), _lib.ptr(m.indices.astype(np.int32)), 0,
                                  _lib.ptr(m.data.astype(np.float64)), 0)
    assert rc == _lib.HB_ERR_NO_DEVICE and "no CPU fallback" in _lib.last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "hicexplorer_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn + " must not reference oracle/"


def test_build_produces_both_libraries_and_the_io_library_exports_its_codecs():
    """__graft_entry__.build() must leave libhicbalance.so AND libhbio.so in the tree (a fresh checkout otherwise falls
    back to the pure-Python chunk decoders, orders of magnitude slower on real files)"""
    import ctypes
    import os
    from hicexplorer_b200 import _build
    pkg = os.path.dirname(os.path.abspath(_build.__file__))
    assert os.path.exists(os.path.join(pkg, "libhicbalance.so"))
    io_path = os.path.join(pkg, "libhbio.so")
    if not os.path.exists(io_path):
        _build.build_io()
    lib = ctypes.CDLL(io_path)
    for name in ("hbio_blosc_decode", "hbio_decode_chunks", "hbio_unshuffle", "hbio_blosc_encode", "hbio_encode_chunks",
                 "hbio_encode_bound", "hbio_blosc_bound"):
        assert hasattr(lib, name), name
    # and the build function itself reaches build_io(): the `out` parameter is not shadowed any more
    import inspect
    src = inspect.getsource(_build.build)
    assert "out = p.communicate" not in src and "build_io()" in src
