"""Time hb_csr_create on a host CSR of raw counts, pinned and pageable, with the narrow ingest on (default) or off
(HB_NARROW_INGEST=0 in the environment): python tools/ingest_probe.py [entries, default 4e8]"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hicexplorer_b200 import _lib  # noqa: E402

L = _lib.lib()
nnz = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
n = 100_000
per_row = nnz // n
nnz = per_row * n
res = {"entries": nnz, "narrow": os.environ.get("HB_NARROW_INGEST", "1") != "0"}
for kind in ("pinned", "pageable"):
    pin = kind == "pinned"
    ip = torch.arange(0, nnz + 1, per_row, dtype=torch.int64)
    ix = torch.empty(nnz, dtype=torch.int32, pin_memory=pin)
    da = torch.empty(nnz, dtype=torch.float64, pin_memory=pin)
    cols = torch.arange(per_row, dtype=torch.int32)
    ixv, dav = ix.view(n, per_row), da.view(n, per_row)
    ixv[:] = cols[None, :]
    ixv += (torch.arange(n, dtype=torch.int32) % (n - per_row))[:, None]
    dav[:] = (cols % 1000).to(torch.float64)[None, :]
    for rep in range(3):
        h = ctypes.c_void_p()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.hb_csr_create(ctypes.byref(h), n, n, 0, nnz, ctypes.c_void_p(ip.data_ptr()), ctypes.c_void_p(ix.data_ptr()), 0,
                                   ctypes.c_void_p(da.data_ptr()), 0))
        dt = time.perf_counter() - t0
        k, b = ctypes.c_int(), ctypes.c_int64()
        L.hb_csr_ingest_info(h, ctypes.byref(k), ctypes.byref(b))
        L.hb_csr_destroy(h)
        res["%s_%d" % (kind, rep)] = {"create_s": round(dt, 4), "host_GBps": round(12.0 * nnz / dt / 1e9, 1), "kind": k.value,
                                      "pcie_GB": round(b.value / 1e9, 2)}
    del ix, da, ixv, dav
print(json.dumps(res))
