#!/usr/bin/env python
"""profiles/r2_traffic.json from an `ncu --set full` raw CSV export of hb_ice_loop_kernel:
   ncu -i gpurun_out/X.ncu-rep --page raw --csv > X.raw.csv ; python tools/ncu_traffic.py X.raw.csv ice10k ITERATIONS
bench.py reports the entry as `roofline.traffic` (DRAM bytes per iteration) as long as the kernel sources are the ones
the capture was taken on (source_sha256 over hb_spmv.cuh + hb_ice.cu + hb_internal.cuh)."""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, workload, iterations = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def num(key):
        v, u = d[key]
        x = float(v.replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
                 "s": 1.0, "%": 1.0}.get(u, 1.0)
        return x * scale

    sys.path.insert(0, ROOT)
    import bench
    method, bins, kw = bench.synth_kwargs(workload, clean=True)
    src = b"".join(open(os.path.join(ROOT, "hicexplorer_b200", "csrc", f), "rb").read()
                   for f in ("hb_spmv.cuh", "hb_ice.cu", "hb_internal.cuh"))
    out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    nnz = int(sys.argv[4]) if len(sys.argv) > 4 else None
    out[workload] = {
        "kernel": d["Kernel Name"][0], "iterations": iterations, "nnz": nnz,
        "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
        "gpu_time_s": num("gpu__time_duration.sum"),
        "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "source_sha256": hashlib.sha256(src).hexdigest(), "from": os.path.basename(path)}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out[workload], indent=1))


if __name__ == "__main__":
    main()
