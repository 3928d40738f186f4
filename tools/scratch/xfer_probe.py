"""D2H / H2D staged-copy sweep: workers x chunk, fresh vs pre-faulted pageable destination."""
import os, sys, time, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.getcwd())
    import numpy as np
    import bench
    from hicexplorer_b200.store import DeviceCSR
    method, bins, kw = bench.synth_kwargs("ice25k", clean=True)
    bins = bins[:12]
    with DeviceCSR.synthetic(bins, **kw) as dev:
        nnz = dev.nnz
        res = {}
        for rep in range(2):
            t0 = time.time(); m = dev.download(); t = time.time() - t0
            res["d2h_fresh_GBps_%d" % rep] = 12 * nnz / t / 1e9
            t0 = time.time(); m2 = dev.download(); t = time.time() - t0
            del m2
        import ctypes
        from hicexplorer_b200 import _lib
        t0 = time.time()
        _lib.check(_lib.lib().hb_csr_download(dev.handle, None, _lib.ptr(m.indices), _lib.ptr(m.data)))
        res["d2h_prefaulted_GBps"] = 12 * nnz / (time.time() - t0) / 1e9
    t0 = time.time(); d = DeviceCSR.from_scipy(m); t = time.time() - t0; d.close()
    res["h2d_GBps"] = 12 * nnz / t / 1e9
    res["GB"] = 12 * nnz / 1e9
    print(json.dumps(res))
else:
    print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
    for w, c in [(6, 16), (12, 16), (16, 16), (12, 4), (12, 64), (3, 16)]:
        env = dict(os.environ, HB_XFER_WORKERS=str(w), HB_XFER_CHUNK_MB=str(c))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(w, c, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:], flush=True)
