#!/usr/bin/env python
"""Wall time of the whole drop-in command on a genome-wide file, as a user runs it:
    hicCorrectMatrix correct -m <synthetic 10 kb .h5> -o out.h5 --correctionMethod ICE --filterThreshold -0.85 50
The input file is written first (synthetic matrix generated on the device, upper triangle, this repo's writer); the
command then runs in-process under cProfile so that the host-side stages show up next to the device time.
Usage: python tools/cli_e2e.py [workload] [--cool] [--kr]      Prints one JSON line.  (--kr: the default method instead)"""
import cProfile
import io
import json
import os
import pstats
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from hicexplorer_b200 import hicCorrectMatrix
    from hicexplorer_b200.hicmatrix_lite import hiCMatrix
    from hicexplorer_b200.store import DeviceCSR
    workload = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "ice10k"
    ext = ".cool" if "--cool" in sys.argv else ".h5"
    method, bins, kw = bench.synth_kwargs(workload, clean=False)
    res = bench.WORKLOADS[workload][1]
    tmp = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))
    src, dst = os.path.join(tmp, "in" + ext), os.path.join(tmp, "out" + ext)
    t0 = time.time()
    with DeviceCSR.synthetic(bins, **kw) as dev:
        dev.keep_upper()
        up = dev.download()
    ma = hiCMatrix()
    ma._set_upper(up)
    ma.cut_intervals = [("chr%d" % (c + 1), i * res, (i + 1) * res, 1.0) for c, L in enumerate(bins) for i in range(L)]
    ma._index_intervals()
    ma.value_dtype = np.int32
    ma._file_layout = up
    ma.save(src)
    del ma, up
    t_prepare = time.time() - t0
    prof = cProfile.Profile()
    t0 = time.time()
    prof.enable()
    if "--kr" in sys.argv:
        hicCorrectMatrix.main(["correct", "-m", src, "-o", dst, "--correctionMethod", "KR"])
    else:
        hicCorrectMatrix.main(["correct", "-m", src, "-o", dst, "--correctionMethod", "ICE", "--filterThreshold", "-0.85", "50"])
    prof.disable()
    t_cmd = time.time() - t0
    buf = io.StringIO()
    pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(28)
    top = [ln.strip() for ln in buf.getvalue().splitlines() if ln.strip() and ln.lstrip()[0].isdigit()][:28]
    out = {"workload": bench.workload_name(workload, bins, kw), "file": ext, "method": "KR" if "--kr" in sys.argv else "ICE", "input_bytes": os.path.getsize(src),
           "output_bytes": os.path.getsize(dst), "prepare_input_s": round(t_prepare, 2), "command_s": round(t_cmd, 2),
           "profile_top_cumulative": top}
    print(json.dumps(out))
    os.remove(src)
    os.remove(dst)


if __name__ == "__main__":
    main()
