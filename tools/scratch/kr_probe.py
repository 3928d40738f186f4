import os, sys
sys.path.insert(0, os.getcwd())
import bench
from hicexplorer_b200.store import DeviceCSR
method, bins, kw = bench.synth_kwargs("kr25k", clean=False)
with DeviceCSR.synthetic(bins, **kw) as dev:
    x, outer, mv, res = dev.kr()
    print("KR", outer, mv, res)
