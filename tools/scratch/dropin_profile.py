import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from hicexplorer_b200.store import DeviceCSR
from hicexplorer_b200.iterativeCorrection import iterativeCorrection
wl = sys.argv[1] if len(sys.argv) > 1 else "ice10k"
method, bins, kw = bench.synth_kwargs(wl, clean=True)
with DeviceCSR.synthetic(bins, **kw) as dev:
    m = dev.download()
prof = cProfile.Profile()
t0 = time.time()
prof.enable()
w, b = iterativeCorrection(m, M=200, tolerance=0.0)
prof.disable()
print("total %.2f s" % (time.time() - t0))
buf = io.StringIO()
pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(22)
print("\n".join(l[:150] for l in buf.getvalue().splitlines() if l.strip() and l.lstrip()[0].isdigit()))
