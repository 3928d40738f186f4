"""Device symmetrisation (hb_csr_symmetrize_upper) of the synthetic 10 kb matrix: rank ordering vs the round-1 sorting
network (HB_SYM_SORT=network).  python tools/sym_timing.py [workload]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hicexplorer_b200.store import DeviceCSR  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "ice10k"
method, bins, kw = bench.synth_kwargs(workload, clean=True)
out = {"workload": workload}
for mode in ("rank", "network", "rank", "network"):
    if mode == "network":
        os.environ["HB_SYM_SORT"] = "network"
    else:
        os.environ.pop("HB_SYM_SORT", None)
    with DeviceCSR.synthetic(bins, **kw) as dev:
        dev.keep_upper()
        dev.sync()
        t0 = time.perf_counter()
        dev.symmetrize_upper()
        dev.sync()
        dt = time.perf_counter() - t0
        out.setdefault(mode + "_ms", []).append(round(dt * 1e3, 2))
        out["nnz_full"] = dev.nnz
        out.setdefault("symmetry_ratio", []).append(dev.symmetry_ratio())
print(json.dumps(out))
