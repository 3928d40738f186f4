#!/usr/bin/env python
"""Fixed cost of one ICE iteration / one KR mat-vec inside the persistent kernels, measured on workloads of different
size on one GPU: ms per iteration of hb_dev_ice_loop, of the step-level path, and of the stand-alone SpMV kernel.
   python tools/loop_overhead.py [workload ...] > gpurun_out/loop_overhead.json"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from hicexplorer_b200 import _lib
    from hicexplorer_b200.store import DeviceCSR
    L = _lib.lib()
    check = _lib.check
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    out = {}
    for wl in (sys.argv[1:] or ["ice100k", "ice25k", "ice10k"]):
        method, bins, kw = bench.synth_kwargs(wl, clean=True)
        with DeviceCSR.synthetic(bins, **kw) as dev:
            h = dev.handle
            n, nnz = dev.n, dev.nnz
            res = {"bins": n, "nnz": nnz}
            for packed in (False, True):
                if packed:
                    dev.pack_values()
                tag = "packed_" if packed else ""
                K = 50

                def timed(fn):
                    check(L.hb_dev_ice_begin(h, sptr))
                    fn(5)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    fn(K)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    return e0.elapsed_time(e1) / K

                def loop(k):
                    check(L.hb_dev_ice_loop(h, k, 0.0, sptr))

                def steps(k):
                    for _ in range(k):
                        check(L.hb_dev_ice_spmv(h, None, 0, 1, sptr))
                        check(L.hb_dev_ice_update(h, None, 1, 0.0, sptr))

                def spmv_only(k):
                    for _ in range(k):
                        check(L.hb_dev_ice_spmv(h, None, 0, 1, sptr))

                res[tag + "loop_ms"] = timed(loop)
                res[tag + "step_level_ms"] = timed(steps)
                res[tag + "spmv_only_ms"] = timed(spmv_only)
                t0 = time.time()
                x, outer, mv, r = dev.kr()
                t0 = time.time()
                x, outer, mv, r = dev.kr()
                kr_s = time.time() - t0
                res[tag + "kr_ms_per_matvec"] = kr_s * 1e3 / (mv + 1)
                res[tag + "kr_matvecs"] = mv
                os.environ["HB_NO_PERSISTENT"] = "1"
                t0 = time.time()
                dev.kr()
                res[tag + "kr_step_level_ms_per_matvec"] = (time.time() - t0) * 1e3 / (mv + 1)
                del os.environ["HB_NO_PERSISTENT"]
            out[wl] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
