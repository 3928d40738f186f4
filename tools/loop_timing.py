#!/usr/bin/env python
"""Where one iteration of hb_ice_loop_kernel spends its time (block 0's %globaltimer stamps; HB_LOOP_TIMING=1).
   python tools/loop_timing.py [workload ...]   -> one line per workload on stderr"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["HB_LOOP_TIMING"] = "1"
import bench  # noqa: E402


def main():
    from hicexplorer_b200 import _lib
    from hicexplorer_b200.store import DeviceCSR
    L = _lib.lib()
    for wl in (sys.argv[1:] or ["ice100k", "ice25k", "ice10k"]):
        method, bins, kw = bench.synth_kwargs(wl, clean=True)
        with DeviceCSR.synthetic(bins, **kw) as dev:
            for packed in (False, True):
                if packed:
                    dev.pack_values()
                sys.stderr.write("%s%s: " % (wl, " packed" if packed else ""))
                sys.stderr.flush()
                _lib.check(L.hb_dev_ice_begin(dev.handle, None))
                _lib.check(L.hb_dev_ice_loop(dev.handle, 24, 0.0, None))
                dev.sync()


if __name__ == "__main__":
    main()
