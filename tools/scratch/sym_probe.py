import sys, time
sys.path.insert(0, '/root/repo')
import bench
from hicexplorer_b200.store import DeviceCSR
method, bins, kw = bench.synth_kwargs(sys.argv[1] if len(sys.argv) > 1 else "ice10k", clean=False)
with DeviceCSR.synthetic(bins, **kw) as dev:
    dev.keep_upper()
    up = dev.download()
for rep in range(2):
    t0 = time.time()
    d = DeviceCSR.from_upper(up)
    print("from_upper %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
    d.close()
