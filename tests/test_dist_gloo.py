"""World-size-2 CPU (gloo) tests of the N > 1 host logic: the nnz-balanced row partition, the collective
callback the C library calls (hb_allreduce_fn, exercised here through ctypes with host buffers), and the
exchange scheme itself -- every rank fills only its rows of a zeroed vector, one sum-allreduce assembles it,
the O(n) update then runs identically on all ranks.  The per-rank SpMV is stood in for by scipy here (there
is no GPU in this test); the CUDA version of the same loop is covered by the -m gpu tests and bench.py --gpus N."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from hicexplorer_b200.dist import make_allreduce_callback, nnz_balanced_partition
from oracle import ice_oracle
from tests.helpers import random_hic


def test_partition_is_contiguous_complete_and_nnz_balanced():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 5000, 10000)
    counts[rng.random(10000) < 0.05] = 0
    indptr = np.concatenate([[0], np.cumsum(counts)])
    for world in (1, 2, 3, 4, 8):
        b = nnz_balanced_partition(indptr, world)
        assert b[0] == 0 and b[-1] == 10000 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        share = np.diff(indptr[b])
        assert share.sum() == indptr[-1]
        assert share.max() - share.min() <= 2 * counts.max()     # equal shares up to one row
    # degenerate inputs: empty matrix, more ranks than rows
    assert list(nnz_balanced_partition(np.zeros(1, dtype=np.int64), 4)) == [0, 0, 0, 0, 0]
    b = nnz_balanced_partition(np.array([0, 3, 6]), 4)
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sharded_ice_worker(rank, world, port, n, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = random_hic(n, 0.05, seed, empty_frac=0.02)
        bounds = nnz_balanced_partition(m.indptr, world)
        a, b = int(bounds[rank]), int(bounds[rank + 1])
        block = m[a:b]                                   # this rank's rows, all columns
        fn, keep = make_allreduce_callback(device_buffers=False)
        exch = np.zeros(n + world, dtype=np.float64)
        bias = np.ones(n)
        r = np.ones(n)
        iters = 0
        for _ in range(200):
            exch[:] = 0.0
            exch[a:b] = r[a:b] * (block @ r)             # stand-in for hb_spmv_kernel<HbIceEpi>
            exch[n + rank] = 1.0 + rank                  # per-rank maximum slot
            rc = fn(None, exch.ctypes.data_as(ctypes.c_void_p), n + world, None)   # what libhicbalance calls
            assert rc == 0
            assert np.array_equal(exch[n:], 1.0 + np.arange(world))
            s_raw = exch[:n]
            s = s_raw / s_raw[s_raw != 0].mean()
            bias *= s
            dev = np.abs(s - 1).max()
            live = s_raw != 0
            r[live] = r[live] * (1.0 / s[live])
            r[~live] = 0.0
            iters += 1
            if dev < 1e-5:
                break
        bias /= bias[bias != 0].mean()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), bias=bias, iters=iters, a=a, b=b)
        del keep
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_ice_exchange_scheme_world2(tmp_path):
    n, seed, world = 700, 4, 2
    port = _free_port()
    mp.spawn(_sharded_ice_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    m = random_hic(n, 0.05, seed, empty_frac=0.02)
    ref = ice_oracle.ice(m.copy(), max_iter=200, tolerance=1e-5)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert r0["b"] == r1["a"] and r0["a"] == 0 and r1["b"] == n
    np.testing.assert_array_equal(r0["bias"], r1["bias"])        # replicated state is bit-identical across ranks
    assert abs(int(r0["iters"]) - ref["iterations"]) <= 1
    np.testing.assert_allclose(r0["bias"], ref["bias"], rtol=1e-9)


def test_lpt_assignment_is_deterministic_and_balanced():
    from hicexplorer_b200.dist import lpt_assignment
    w = [248, 242, 198, 190, 181, 170, 159, 145, 138, 133, 135, 133, 114, 107, 101, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    for world in (1, 2, 3, 8):
        owner = lpt_assignment(w, world)
        assert owner == lpt_assignment(list(w), world) and set(owner) == set(range(world))
        loads = [sum(x for x, o in zip(w, owner) if o == r) for r in range(world)]
        assert max(loads) <= sum(w) / world + max(w) * 0.5
