"""The sharded (N > 1) code path of libhicbalance on ONE GPU: `world` row-block handles live in one process, one
host thread per emulated rank runs the same hb_ice_run / hb_kr_run a real rank would, and the collective callback
(hb_allreduce_fn) is implemented with a host barrier + a device-side sum.  No kernel waits on another kernel
(B200_PROFILING.md); only host threads wait.  Property checked: results are BIT-IDENTICAL to the single-block run,
for any number of blocks -- the fixed-tile reductions make the arithmetic independent of the sharding."""
import ctypes
import threading

import numpy as np
import pytest

from hicexplorer_b200 import _lib
from hicexplorer_b200.dist import _DevicePtr, nnz_balanced_partition
from hicexplorer_b200.store import DeviceCSR
from oracle import ice_oracle, kr_oracle
from tests.helpers import full_from_pixels, random_hic

pytestmark = pytest.mark.gpu


class FakeComm(object):
    """sum-allreduce between `world` threads of one process (device buffers on the same GPU)."""

    def __init__(self, world):
        import torch
        self.torch = torch
        self.world = world
        self.barrier = threading.Barrier(world)
        self.bufs = [None] * world
        self.calls = 0

    def callback_for(self, rank):
        torch = self.torch

        def cb(_ctx, d_buf, count, stream):
            try:
                torch.cuda.ExternalStream(stream).synchronize()
                self.bufs[rank] = torch.as_tensor(_DevicePtr(d_buf, count), device="cuda")
                self.barrier.wait()
                if rank == 0:
                    total = self.bufs[0].clone()
                    for t in self.bufs[1:]:
                        total += t
                    for t in self.bufs:
                        t.copy_(total)
                    torch.cuda.synchronize()
                    self.calls += 1
                self.barrier.wait()
                return 0
            except Exception as exc:  # pragma: no cover
                import sys
                sys.stderr.write("fake allreduce failed: %r\n" % (exc,))
                self.barrier.abort()
                return 1

        return _lib.ALLREDUCE_FN(cb)


def _run_sharded(matrix, world, fn):
    bounds = nnz_balanced_partition(matrix.indptr, world)
    comm = FakeComm(world)
    results = [None] * world
    errors = []

    def worker(rank):
        try:
            a, b = int(bounds[rank]), int(bounds[rank + 1])
            with DeviceCSR.from_scipy(matrix, row_range=(a, b)) as dev:
                cb = comm.callback_for(rank)
                _lib.check(_lib.lib().hb_set_comm(dev.handle, rank, world, ctypes.cast(cb, ctypes.c_void_p), None))
                results[rank] = fn(dev)
        except Exception as exc:
            errors.append(exc)
            comm.barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    return results, comm


@pytest.mark.parametrize("world", [2, 3, 5])
def test_sharded_ice_is_bit_identical_to_single_block(world, gm12878):
    g = gm12878
    full = full_from_pixels(g)
    ref = ice_oracle.correct_ice_pipeline(full, -1.5, 5.0, max_iter=1)      # only to get the filtered ICE input
    m = ref["ice_input"]
    m.sort_indices()
    with DeviceCSR.from_scipy(m) as dev:
        bias1, it1, dev1 = dev.ice(500, 1e-5)
    results, comm = _run_sharded(m, world, lambda d: d.ice(500, 1e-5))
    for bias, iters, devn in results:
        assert int(iters[0]) == int(it1[0])
        np.testing.assert_array_equal(bias, bias1)          # bit-identical, replicated on every rank
        assert devn[0] == dev1[0]
    assert comm.calls >= int(it1[0])                        # one exchange per iteration
    np.testing.assert_allclose(bias1, g["ice_bias"], rtol=1e-9)


def test_sharded_ice_with_empty_rows_and_tiny_blocks():
    m = random_hic(300, 0.05, 3, empty_frac=0.05)
    with DeviceCSR.from_scipy(m) as dev:
        bias1, it1, _ = dev.ice(40, 1e-5)
    results, _ = _run_sharded(m, 4, lambda d: d.ice(40, 1e-5))
    for bias, iters, _d in results:
        np.testing.assert_array_equal(bias, bias1)
        assert int(iters[0]) == int(it1[0]) == 40           # empty rows: never converges


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_kr_is_bit_identical_to_single_block(world, gm12878):
    m = full_from_pixels(gm12878)
    with DeviceCSR.from_scipy(m) as dev:
        x1, outer1, mv1, res1 = dev.kr()
    results, comm = _run_sharded(m, world, lambda d: d.kr())
    for x, outer, mv, res in results:
        assert (outer, mv) == (outer1, mv1)
        np.testing.assert_array_equal(x, x1)
    assert comm.calls == mv1 + 1                            # one exchange per mat-vec (+ the initial residual)
    ref = kr_oracle.kr_balance(m)
    np.testing.assert_allclose(x1, ref["x"], rtol=1e-9)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_filter_and_ice_pipeline_matches_single_block(world, gm12878):
    """zero-bin mask -> scrub -> MAD filter -> bin deletion -> ICE, every stage on row blocks."""
    from hicexplorer_b200.pipeline import ice_pipeline
    full = full_from_pixels(gm12878)
    single = ice_pipeline(full, (-1.5, 5.0), max_iter=500, want_matrix=False)
    n = full.shape[0]
    gather = {"rs": np.zeros(n), "lock": threading.Lock(), "barrier": threading.Barrier(world)}

    def stage(dev):
        _n_rows, _n, row0, _nnz = dev.dims()
        rs, _dg = dev.row_stats()                         # block-local rows
        with gather["lock"]:
            gather["rs"][row0:row0 + len(rs)] = rs        # stands in for an all-gather of n doubles on the host
        gather["barrier"].wait()
        zero = gather["rs"] == 0
        dev.compact(zero.astype(np.uint8))
        dev.scrub()
        mask, _z, _med, _mad = dev.mad_filter(-1.5, 5.0)  # one collective inside
        dev.compact(mask)
        bias, iters, devn = dev.ice(500, 1e-5)
        return zero, mask, bias, iters, dev.dims()

    results, _comm = _run_sharded(full, world, stage)
    rows_total = 0
    for zero, mask, bias, iters, dims in results:
        assert np.array_equal(np.flatnonzero(zero), single["zero_bins"])
        assert np.array_equal(np.flatnonzero(mask), single["outliers_rel"])     # bit-exact mask on every rank
        np.testing.assert_array_equal(bias, single["bias"])                     # and bit-identical bias
        assert int(iters[0]) == int(single["iterations"][0])
        assert dims[1] == len(single["kept"])
        rows_total += dims[0]
    assert rows_total == len(single["kept"])                                    # blocks still tile the matrix


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_kr_rescale_and_scaled_upper(world, kr_partial):
    m = full_from_pixels(kr_partial)
    with DeviceCSR.from_scipy(m) as dev:
        x1, _o, _mv, _r = dev.kr()
        xr1, f1 = dev.kr_rescale(x1)
        up1 = dev.kr_scaled_upper(xr1)

    def stage(dev):
        x, _o, _mv, _r = dev.kr()
        xr, f = dev.kr_rescale(x)
        return xr, f, dev.kr_scaled_upper(xr), dev.dims()

    results, _comm = _run_sharded(m, world, stage)
    rows = []
    for xr, f, up, dims in results:
        assert f == f1
        np.testing.assert_array_equal(xr, xr1)
        assert up.shape == (dims[0], m.shape[0])
        rows.append(up)
    from scipy import sparse
    whole = sparse.vstack(rows).tocsr()
    assert np.array_equal(whole.indptr, up1.indptr) and np.array_equal(whole.indices, up1.indices)
    np.testing.assert_array_equal(whole.data, up1.data)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_symmetry_screen(world):
    """row blocks add their exact-symmetry digests with one allreduce: 0 for a symmetric matrix on every rank, a
    rejection (ratio 1) on every rank once a single entry differs from its mirror image"""
    sym = random_hic(700, 0.1, 11)
    assert all(r == 0.0 for r in _run_sharded(sym, world, lambda dev: dev.symmetry_ratio())[0])
    broken = sym.copy()
    k = int(np.flatnonzero(broken.indices > np.repeat(np.arange(700), np.diff(broken.indptr)))[5])
    broken.data[k] += 1e-9
    assert all(r == 1.0 for r in _run_sharded(broken, world, lambda dev: dev.symmetry_ratio())[0])
