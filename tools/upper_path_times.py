#!/usr/bin/env python
"""Timings of the file-layout ends of the pipeline on the synthetic 10 kb matrix (1 GPU): upload of the UPPER triangle
+ symmetrisation in HBM (hb_csr_create_upper) against the upload of the full symmetric matrix (hb_csr_create), the
corrected matrix emitted in file layout (hb_ice_scaled_upper) against the full value array (hb_ice_scaled_values),
and the --transCutoff percentile.  Prints one JSON line.  Usage: python tools/upper_path_times.py [workload]"""
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
    out = {"workload": bench.workload_name(workload, bins, kw)}
    DeviceCSR.synthetic([2000], **kw).close()

    def wall(fn):
        t0 = time.time()
        r = fn()
        return r, (time.time() - t0) * 1e3

    with DeviceCSR.synthetic(bins, **kw) as dev:
        n, nnz = dev.n, dev.nnz
        offs = np.concatenate([[0], np.cumsum(bins)]).astype(np.int64)
        pct, t = wall(lambda: dev.trans_percentile(offs, 100 - 0.0005))
        out["trans_percentile"] = {"ms": round(t, 2), "value": pct, "passes": 9}
        full = None
        if 12 * nnz < 8e9:                      # the full matrix on the host only when it is small (host RAM)
            full, t = wall(dev.download)
            out["download_full_ms"] = round(t, 1)
        dev.keep_upper()
        up, t = wall(dev.download)
        out["download_upper_ms"] = round(t, 1)
    out.update({"bins": n, "nnz_full": nnz, "nnz_upper": int(up.nnz)})
    # upload paths (pageable numpy arrays, like a matrix that came out of a file reader)
    t_full = None
    for rep in range(2):
        if full is not None:
            d, t_full = wall(lambda: DeviceCSR.from_scipy(full))
            d.close()
        d, t_up_only = wall(lambda: DeviceCSR.from_scipy(up))
        d.close()
        d, t_upper = wall(lambda: DeviceCSR.from_upper(up))
        same = None
        if rep == 1:
            out["symmetrised"] = {"nnz": d.nnz, "equals_generated_nnz": d.nnz == nnz, "symmetry_ratio": d.symmetry_ratio()}
        if rep == 1 and full is not None:
            got = d.download()
            same = bool(np.array_equal(got.indptr, full.indptr) and np.array_equal(got.indices, full.indices) and
                        np.array_equal(got.data, full.data))
            del got
        d.close()
    out["upload"] = {"full_symmetric_ms": None if t_full is None else round(t_full, 1), "upper_triangle_only_ms": round(t_up_only, 1),
                     "upper_plus_device_symmetrise_ms": round(t_upper, 1),
                     "device_symmetrise_ms": round(t_upper - t_up_only, 1), "result_identical_to_full_upload": same,
                     "bytes_full": 12 * nnz, "bytes_upper": 12 * int(up.nnz)}
    del full
    with DeviceCSR.from_upper(up) as dev:
        rs, _ = dev.row_stats()
        zero = rs == 0
        dev.compact(zero.astype(np.uint8))
        keep1 = np.flatnonzero(~zero)
        dev.scrub()
        mask = dev.mad_filter(-0.85, 50.0)[0]
        dev.compact(mask)
        kept = keep1[~mask.astype(bool)]
        dev.pack_values()
        (bias, iters, devn), t = wall(lambda: dev.ice(500, 1e-5))
        out["ice"] = {"ms": round(t, 1), "iterations": int(iters[0])}
        vals, t = wall(dev.ice_scaled_values)
        out["emit_full_values_ms"] = round(t, 1)
        nv = len(vals)
        del vals
        upper, t = wall(lambda: dev.ice_scaled_upper(kept, n))
        out["emit_file_layout_ms"] = round(t, 1)
        out["emit"] = {"full_value_bytes": 8 * nv, "file_layout_bytes": 12 * int(upper.nnz) + 8 * (n + 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
