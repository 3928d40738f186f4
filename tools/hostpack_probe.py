"""Host throughput of the narrow-ingest conversion (csrc/hb_hostpack.cpp) on this machine's cores:
python tools/hostpack_probe.py [log2 entries]"""
import ctypes
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
so = os.path.join(tmp, "libhp.so")
subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-o", so,
                       os.path.join(ROOT, "hicexplorer_b200", "csrc", "hb_hostpack.cpp")])
L = ctypes.CDLL(so)
L.hb_host_pack_u16.restype = ctypes.c_int
L.hb_host_pack_u16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
a = np.empty(n, dtype=np.float64)
out = np.empty(n, dtype=np.uint16)
step = 1 << 22
for o in range(0, n, step):
    a[o:o + step] = (np.arange(o, min(o + step, n)) % 60000).astype(np.float64)
out[:] = 0
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for T in (1, 2, 4, 8, 12, 16, 24, 32):
    parts = [(k * n // T, (k + 1) * n // T) for k in range(T)]
    res = [0] * T

    def w(k):
        s, e = parts[k]
        res[k] = L.hb_host_pack_u16(a[s:e].ctypes.data, out[s:e].ctypes.data, e - s)
    best = 1e9
    for _ in range(3):
        ths = [threading.Thread(target=w, args=(k,)) for k in range(T)]
        t = time.perf_counter()
        [x.start() for x in ths]
        [x.join() for x in ths]
        best = min(best, time.perf_counter() - t)
    assert all(res)
    print("%2d threads: %.2f G entries/s, %.1f GB/s read" % (T, n / best / 1e9, 8 * n / best / 1e9), flush=True)
