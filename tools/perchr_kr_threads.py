#!/usr/bin/env python
"""--perchr KR on the synthetic 5 kb matrix (24 independent chromosome blocks, 1 GPU): chromosome by chromosome against
T host threads that each run whole chromosomes on their own handle/stream (the C ABI releases the GIL), which hides the
per-mat-vec host synchronisation of one chromosome behind the kernels of another.  Prints one JSON line."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
bench.WORKLOADS.update({"perchr5k": ("ICE+KR", 5000, 2.0e9, 16000)})


def main():
    from hicexplorer_b200.store import DeviceCSR
    method, bins, kw = bench.synth_kwargs("perchr5k", clean=True)
    kw = dict(kw, n_trans=0)
    with DeviceCSR.synthetic([2000], **kw) as d0:
        d0.kr()
    t0 = time.time()
    devs = [DeviceCSR.synthetic([b], **kw) for b in bins]
    for d in devs:
        d.sync()
    t_gen = time.time() - t0
    nnz = [d.nnz for d in devs]
    out = {"chromosomes": len(bins), "entries": int(sum(nnz)), "generation_ms": round(t_gen * 1e3, 1), "runs": []}
    ref = None
    for threads in (1, 2, 3, 4, 6):
        order = sorted(range(len(bins)), key=lambda c: -nnz[c])
        t0 = time.time()
        if threads == 1:
            res = {c: devs[c].kr() for c in order}
        else:
            with ThreadPoolExecutor(threads) as pool:
                res = dict(zip(order, pool.map(lambda c: devs[c].kr(), order)))
        dt = time.time() - t0
        mv = sum(r[2] + 1 for r in res.values())
        work = sum(float(nnz[c]) * (res[c][2] + 1) for c in res)
        if ref is None:
            ref = res
        same = all(np.array_equal(res[c][0], ref[c][0]) for c in res)
        out["runs"].append({"threads": threads, "ms": round(dt * 1e3, 1), "matvecs": mv, "entries_per_s": work / dt,
                            "GBps": 12.0 * work / dt / 1e9, "bit_identical_to_serial": same})
    for d in devs:
        d.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
