#!/usr/bin/env python
"""Wall time of every device stage of the ICE pipeline on the synthetic 10 kb matrix WITH the planted empty and
low-coverage bins (1 GPU): what `hicCorrectMatrix correct --correctionMethod ICE` does between upload and download.
Prints one JSON line.  Usage: python tools/stage_times.py [workload]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from hicexplorer_b200.store import DeviceCSR
    workload = sys.argv[1] if len(sys.argv) > 1 else "ice10k"
    method, bins, kw = bench.synth_kwargs(workload, clean=False)
    peak, _ = bench.measured_peak()
    stages = []

    def timed(name, fn, bytes_moved=None):
        dev.sync()
        t0 = time.time()
        out = fn()
        dev.sync()
        dt = time.time() - t0
        rec = {"stage": name, "ms": round(dt * 1e3, 3)}
        if bytes_moved:
            rec["GBps"] = round(bytes_moved / dt / 1e9, 1)
            rec["frac_of_peak"] = round(bytes_moved / dt / 1e9 / peak, 3)
        stages.append(rec)
        return out

    with DeviceCSR.synthetic(bins, **kw) as dev:
        DeviceCSR.synthetic([2000], **kw).close()      # warm the lazily loaded kernels
        n, nnz = dev.n, dev.nnz
        rs, dg = timed("row sums + diagonal (zero-bin mask)", dev.row_stats, 12.0 * nnz)
        zero = rs == 0
        timed("delete zero bins (row+column compaction)", lambda: dev.compact(zero.astype(np.uint8)), 2 * 12.0 * nnz + 12.0 * nnz)
        nnz1 = dev.nnz
        timed("NaN/Inf scrub", dev.scrub, 8.0 * nnz1)
        mask = timed("MAD filter (row stats, 2 exact medians, z-scores)", lambda: dev.mad_filter(-0.85, 50.0)[0], 12.0 * nnz1)
        timed("delete outlier bins (compaction)", lambda: dev.compact(mask), 2 * 12.0 * nnz1 + 12.0 * nnz1)
        nnz2, n2 = dev.nnz, dev.n
        ratio = timed("symmetry test", dev.symmetry_ratio, 12.0 * nnz2)
        kind = timed("pack values (lossless uint16 copy)", dev.pack_values, 2 * 8.0 * nnz2 + 2.0 * nnz2)
        bias, iters, devn = timed("ICE to tolerance 1e-5", lambda: dev.ice(500, 1e-5))
        stages[-1]["iterations"] = int(iters[0])
        stages[-1]["ms_per_iteration"] = round(stages[-1]["ms"] / max(int(iters[0]), 1), 3)
        vals = timed("emit corrected values + download (8 B/entry D2H)", dev.ice_scaled_values)
    out = {"workload": bench.workload_name(workload, bins, kw) + " incl. 2% low-coverage and 0.5% empty bins",
           "bins": n, "nnz": nnz, "bins_after_filter": n2, "nnz_after_filter": nnz2, "symmetry_ratio": ratio,
           "pack_kind": kind, "converged": bool(devn[0] < 1e-5), "stages": stages,
           "total_ms": round(sum(s["ms"] for s in stages), 1)}
    print(json.dumps(out))
    del vals


if __name__ == "__main__":
    main()
