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
