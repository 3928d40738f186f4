"""Kernel-only scaling of the SpMV with the block size on ONE GPU (no exchange): time per launch for the full 10 kb
matrix and for 1/2, 1/4, 1/8 row blocks of it.  Separates launch ramp / tail effects from the peer exchange."""
import ctypes, os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import bench
from hicexplorer_b200 import _lib
from hicexplorer_b200.store import DeviceCSR
L = _lib.lib(); check = _lib.check
method, bins, kw = bench.synth_kwargs("ice10k", clean=True)
n = int(sum(bins))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sptr = ctypes.c_void_p(stream.cuda_stream)
out = []
for frac in (1, 2, 4, 8):
    rows = n // frac
    a = (n - rows) // 2                      # a block from the middle of the matrix
    dev = DeviceCSR.synthetic(bins, row_range=(a, a + rows), **kw)
    h = dev.handle
    exch = torch.zeros(n + 64, dtype=torch.float64, device="cuda")
    eptr = ctypes.c_void_p(exch.data_ptr())
    check(L.hb_dev_ice_begin(h, sptr))
    for _ in range(20):
        check(L.hb_dev_ice_spmv(h, eptr, 0, 1, sptr))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 300
    e0.record(stream)
    for _ in range(reps):
        check(L.hb_dev_ice_spmv(h, eptr, 0, 1, sptr))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out.append({"block": "1/%d" % frac, "rows": rows, "nnz": dev.nnz, "ms": ms, "GBps": 12.0 * dev.nnz / ms / 1e6})
    dev.close()
base = out[0]
for o in out:
    o["ideal_ms_from_full"] = base["ms"] * o["nnz"] / base["nnz"]
    o["overhead_us"] = (o["ms"] - o["ideal_ms_from_full"]) * 1e3
print(json.dumps(out))
