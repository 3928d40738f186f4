"""Device -> pageable host memory through hb_copy_d2h (the way corrected values come back to a numpy array):
fresh np.empty vs a buffer whose pages exist already, for a few worker counts (HB_XFER_WORKERS is read once per
process: run once per setting).  python tools/d2h_probe.py [entries]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hicexplorer_b200 import _lib  # noqa: E402
from hicexplorer_b200._lib import check, lib, ptr  # noqa: E402
from hicexplorer_b200.store import DeviceCSR  # noqa: E402

entries = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
bins = [int(entries // 4000)]
out = {"workers": os.environ.get("HB_XFER_WORKERS", "default"), "thp": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()}
with DeviceCSR.synthetic(bins, seed=1, band=2500, d0=400.0) as dev:
    nnz = dev.nnz
    out["entries"] = nnz
    for rep in range(2):
        t0 = time.perf_counter()
        data = np.empty(nnz, dtype=np.float64)
        check(lib().hb_csr_download(dev.handle, None, None, ptr(data)))
        dt = time.perf_counter() - t0
        out.setdefault("fresh_GBps", []).append(round(8.0 * nnz / dt / 1e9, 1))
        t0 = time.perf_counter()
        check(lib().hb_csr_download(dev.handle, None, None, ptr(data)))
        dt = time.perf_counter() - t0
        out.setdefault("touched_GBps", []).append(round(8.0 * nnz / dt / 1e9, 1))
        del data
        import mmap
        t0 = time.perf_counter()
        mm = mmap.mmap(-1, 8 * nnz)
        mm.madvise(mmap.MADV_HUGEPAGE)
        data = np.frombuffer(mm, dtype=np.float64)
        check(lib().hb_csr_download(dev.handle, None, None, ptr(data)))
        dt = time.perf_counter() - t0
        out.setdefault("fresh_hugepage_GBps", []).append(round(8.0 * nnz / dt / 1e9, 1))
        del data, mm
print(json.dumps(out))
