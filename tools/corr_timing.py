"""Time hicTransform --method pearson on one synthetic chromosome block: device path (densify + centre + cuBLAS DGEMM +
scale + non-zero extraction + D2H) against np.corrcoef on the host cores.  python tools/corr_timing.py [bins]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hicexplorer_b200.store import DeviceCSR  # noqa: E402


def main():
    bins = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
    with DeviceCSR.synthetic([bins], seed=3, band=min(bins, 2000)) as dev:
        m = dev.download()
        dev.block_correlation(0, min(bins, 512), pearson=True)      # opens cuBLAS, first-use costs
        t0 = time.perf_counter()
        got = dev.block_correlation(0, bins, pearson=True)
        t_gpu = time.perf_counter() - t0
    dense = m.toarray()
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        ref = np.nan_to_num(np.corrcoef(dense), nan=0.0, posinf=0.0, neginf=0.0)
    t_cpu = time.perf_counter() - t0
    err = float(np.abs(got.toarray() - ref).max())
    print(json.dumps({"bins": bins, "nnz_in": int(m.nnz), "nnz_out": int(got.nnz), "gpu_s": t_gpu, "numpy_s": t_cpu,
                      "host_threads": os.cpu_count(), "max_abs_diff": err,
                      "dgemm_tflops_if_all_gpu_time": 2.0 * bins ** 3 / t_gpu / 1e12}))


if __name__ == "__main__":
    main()
