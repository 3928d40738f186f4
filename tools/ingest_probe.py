"""Time hb_csr_create on a host CSR of raw counts (pinned and pageable), narrow ingest on and off in the same process.
One GPU:  python tools/ingest_probe.py [total entries, default 4e8]
N GPUs:   python -m torch.distributed.run --nproc-per-node N ... tools/ingest_probe.py 1.5e9   (entries are split over the
ranks; every rank uploads its share at the same time; times are the maximum over the ranks)"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hicexplorer_b200 import _lib  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


L = _lib.lib()
total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["pinned", "pageable"]
n = 100_000
per_row = total // world // n
nnz = per_row * n
res = {"entries_total": nnz * world, "gpus": world, "host_threads": os.cpu_count()}
for kind in kinds:
    pin = kind == "pinned"
    ip = torch.arange(0, nnz + 1, per_row, dtype=torch.int64)
    ix = torch.empty(nnz, dtype=torch.int32, pin_memory=pin)
    da = torch.empty(nnz, dtype=torch.float64, pin_memory=pin)
    cols = torch.arange(per_row, dtype=torch.int32)
    ixv, dav = ix.view(n, per_row), da.view(n, per_row)
    ixv[:] = cols[None, :]
    ixv += (torch.arange(n, dtype=torch.int32) % (n - per_row))[:, None]
    dav[:] = (cols % 1000).to(torch.float64)[None, :]
    modes = [("1", "1"), ("1", "0"), ("0", "0")] * 2       # (narrow side, direct feeder forced on / off): hybrid, narrow only, plain copy
    for narrow, direct in modes:
        os.environ["HB_NARROW_INGEST"] = narrow
        os.environ["HB_INGEST_DIRECT"] = direct
        times = []
        for rep in range(3):
            h = ctypes.c_void_p()
            barrier()
            t0 = time.perf_counter()
            _lib.check(L.hb_csr_create(ctypes.byref(h), n, n, 0, nnz, ctypes.c_void_p(ip.data_ptr()), ctypes.c_void_p(ix.data_ptr()), 0,
                                       ctypes.c_void_p(da.data_ptr()), local))
            barrier()
            times.append(max_over_ranks(time.perf_counter() - t0))
            k, b = ctypes.c_int(), ctypes.c_int64()
            L.hb_csr_ingest_info(h, ctypes.byref(k), ctypes.byref(b))
            L.hb_csr_destroy(h)
        res.setdefault(kind, []).append({"mode": {"11": "hybrid", "10": "narrow only", "00": "plain copy"}[narrow + direct],
                                         "workers": os.environ.get("HB_INGEST_WORKERS", "auto"),
                                         "create_s": [round(t, 4) for t in times], "layout": k.value,
                                         "pcie_GB_per_gpu": round(b.value / 1e9, 2),
                                         "host_GBps_aggregate": round(12.0 * nnz * world / min(times) / 1e9, 1)})
    del ix, da, ixv, dav
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
