"""The persistent solver kernels (hb_ice_loop_kernel / hb_kr_solve_kernel: the whole ICE loop / the whole Knight-Ruiz
solve in ONE cooperative launch, grid barriers instead of kernel boundaries, nothing read back by the host) against
the step-level path they replace (one launch per operation, host in the loop; still used with a host collective and
selectable with HB_NO_PERSISTENT=1).  Both run the same element code and the same fixed-order reductions, so the
property checked everywhere is BIT identity -- plus the oracle on top.

The peer-store exchange (rows stored straight into every rank's replica from the SpMV epilogue, arrival flags
instead of a collective) is covered on ONE GPU by running `world` row-block handles of one process concurrently, one
host thread per rank (hb_peer_attach_inprocess): their persistent kernels are co-resident and wait for each other's
flags exactly as they do across GPUs."""
import contextlib
import ctypes
import os
import threading

import numpy as np
import pytest

from hicexplorer_b200 import _lib
from hicexplorer_b200.dist import attach_peer_inprocess, nnz_balanced_partition
from hicexplorer_b200.store import DeviceCSR
from oracle import ice_oracle, kr_oracle
from tests.helpers import full_from_pixels, random_hic

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def step_level():
    old = os.environ.get("HB_NO_PERSISTENT")
    os.environ["HB_NO_PERSISTENT"] = "1"
    try:
        yield
    finally:
        if old is None:
            del os.environ["HB_NO_PERSISTENT"]
        else:
            os.environ["HB_NO_PERSISTENT"] = old


def _ice_input(gm12878):
    full = full_from_pixels(gm12878)
    m = ice_oracle.correct_ice_pipeline(full, -1.5, 5.0, max_iter=1)["ice_input"]
    m.sort_indices()
    return m


@pytest.mark.parametrize("pack", [False, True])
def test_ice_loop_kernel_equals_step_level_path(gm12878, pack):
    m = _ice_input(gm12878)
    with DeviceCSR.from_scipy(m) as dev:
        if pack:
            assert dev.pack_values() in (1, 2)      # counts above 65535 in this matrix -> float32 copy
        launches0 = _lib.lib().hb_launch_count()
        bias, iters, devn = dev.ice(500, 1e-5)
        launches = _lib.lib().hb_launch_count() - launches0
        with step_level():
            bias_s, iters_s, devn_s = dev.ice(500, 1e-5)
    assert int(iters[0]) == int(iters_s[0]) == 35 or abs(int(iters[0]) - 35) <= 1
    assert int(iters[0]) == int(iters_s[0])
    np.testing.assert_array_equal(bias, bias_s)
    assert devn[0] == devn_s[0]
    assert launches <= 6                     # begin + ONE loop launch + finish (2) -- not 3 launches per iteration
    np.testing.assert_allclose(bias, gm12878["ice_bias"], rtol=1e-9)


def test_ice_loop_kernel_never_converging_and_zero_iterations():
    m = random_hic(400, 0.05, 5, empty_frac=0.05)      # empty rows pin the deviation at >= 1 (iterativeCorrection.py:47)
    with DeviceCSR.from_scipy(m) as dev:
        bias, iters, devn = dev.ice(37, 1e-5)
        with step_level():
            bias_s, iters_s, devn_s = dev.ice(37, 1e-5)
        b0, it0, _ = dev.ice(0, 1e-5)
    assert int(iters[0]) == int(iters_s[0]) == 37
    np.testing.assert_array_equal(bias, bias_s)
    assert devn[0] == devn_s[0] and devn[0] >= 1.0
    assert int(it0[0]) == 0
    ref = ice_oracle.ice(m.copy(), max_iter=37, tolerance=1e-5)
    np.testing.assert_allclose(bias, ref["bias"], rtol=1e-9)


def test_ice_loop_kernel_with_segments():
    """--perchr: segments converge independently inside the one launch"""
    rng = np.random.default_rng(3)
    from scipy import sparse
    blocks = [random_hic(int(k), 0.2, 10 + i) for i, k in enumerate(rng.integers(80, 700, 7))]
    m = sparse.block_diag(blocks, format="csr")
    m.sort_indices()
    offs = np.concatenate([[0], np.cumsum([b.shape[0] for b in blocks])]).astype(np.int64)
    with DeviceCSR.from_scipy(m) as dev:
        dev.set_segments(offs)
        bias, iters, devn = dev.ice(300, 1e-5)
        with step_level():
            bias_s, iters_s, devn_s = dev.ice(300, 1e-5)
    np.testing.assert_array_equal(iters, iters_s)
    np.testing.assert_array_equal(bias, bias_s)
    np.testing.assert_array_equal(devn, devn_s)
    assert len(set(int(i) for i in iters)) > 1          # they did stop at different iterations
    for k, blk in enumerate(blocks):
        ref = ice_oracle.ice(blk.copy(), max_iter=300, tolerance=1e-5)
        assert abs(int(iters[k]) - ref["iterations"]) <= 1
        np.testing.assert_allclose(bias[offs[k]:offs[k + 1]], ref["bias"], rtol=1e-9)


def test_ice_loop_in_chunks_equals_single_steps(gm12878):
    """hb_dev_ice_loop continues from the current state: 3 + 5 iterations == 8 x (spmv, update)"""
    L = _lib.lib()
    m = _ice_input(gm12878)
    n = m.shape[0]
    out = []
    for chunks in ([3, 5], None):
        with DeviceCSR.from_scipy(m) as dev:
            h = dev.handle
            _lib.check(L.hb_dev_ice_begin(h, None))
            if chunks:
                for c in chunks:
                    _lib.check(L.hb_dev_ice_loop(h, c, 0.0, None))
            else:
                for _ in range(8):
                    _lib.check(L.hb_dev_ice_spmv(h, None, 0, 1, None))
                    _lib.check(L.hb_dev_ice_update(h, None, 1, 0.0, None))
            dev.sync()
            bias = np.empty(n)
            iters = np.zeros(1, dtype=np.int32)
            devn = np.zeros(1)
            _lib.check(L.hb_ice_results(h, _lib.ptr(bias), _lib.ptr(iters), _lib.ptr(devn)))
            out.append((bias, int(iters[0]), devn[0]))
    assert out[0][1] == out[1][1] == 8
    np.testing.assert_array_equal(out[0][0], out[1][0])
    assert out[0][2] == out[1][2]


def test_ice_loop_overflow_guard():
    """a value above 1e100 stops the launch and surfaces as the reference's exit (iterativeCorrection.py:58-62)"""
    from scipy import sparse
    m = sparse.csr_matrix(np.array([[0, 1e-200, 0], [1e-200, 0, 1.0], [0, 1.0, 1e250]]))
    with DeviceCSR.from_scipy(m) as dev:
        with pytest.raises(_lib.OverflowIter):
            dev.ice(50, 1e-5)


@pytest.mark.parametrize("name", ["gm12878", "kr_partial", "random"])
def test_kr_solve_kernel_equals_step_level_path(name, gm12878, kr_partial):
    m = {"gm12878": lambda: full_from_pixels(gm12878), "kr_partial": lambda: full_from_pixels(kr_partial),
         "random": lambda: random_hic(900, 0.05, 21, low_frac=0.05)}[name]()
    with DeviceCSR.from_scipy(m) as dev:
        dev.pack_values()
        launches0 = _lib.lib().hb_launch_count()
        x, outer, mv, res = dev.kr()
        launches = _lib.lib().hb_launch_count() - launches0
        with step_level():
            xs, outer_s, mv_s, res_s = dev.kr()
    assert (outer, mv) == (outer_s, mv_s)
    np.testing.assert_array_equal(x, xs)
    assert res == res_s
    assert launches == 1                       # the whole solve: one launch
    ref = kr_oracle.kr_balance(m)
    assert mv == ref["matvecs"] and outer == ref["outer"]
    np.testing.assert_allclose(x, ref["x"], rtol=1e-9)


def test_kr_solve_kernel_cone_boundary_branches():
    """a badly scaled matrix drives the inner CG into the delta / Delta step limiters (bnewt's gamma branches)"""
    rng = np.random.default_rng(8)
    m = random_hic(500, 0.1, 8)
    scale = np.exp(3.0 * rng.standard_normal(500))
    from scipy import sparse
    d = sparse.diags(scale)
    m = sparse.csr_matrix(d @ m @ d)
    m.sort_indices()
    with DeviceCSR.from_scipy(m) as dev:
        x, outer, mv, res = dev.kr()
        with step_level():
            xs, outer_s, mv_s, res_s = dev.kr()
    assert (outer, mv) == (outer_s, mv_s)
    np.testing.assert_array_equal(x, xs)
    ref = kr_oracle.kr_balance(m)
    assert ref["boundary_steps"] > 0           # the branch was taken
    assert mv == ref["matvecs"]
    np.testing.assert_allclose(x, ref["x"], rtol=1e-8)


# ---------------------------------------------------------------------------------------------------------------
def _run_peer(matrix, world, fn, persistent=True):
    """`world` row blocks of one process on ONE GPU in peer mode; one host thread per rank"""
    bounds = nnz_balanced_partition(matrix.indptr, world)
    devs = [DeviceCSR.from_scipy(matrix, row_range=(int(bounds[r]), int(bounds[r + 1]))) for r in range(world)]
    results = [None] * world
    errors = []
    try:
        attach_peer_inprocess(devs)
        for d in devs:
            assert _lib.lib().hb_peer_active(d.handle) == 1

        def worker(rank):
            try:
                results[rank] = fn(devs[rank])
            except Exception as exc:
                errors.append(exc)

        ctx = contextlib.nullcontext() if persistent else step_level()
        with ctx:
            threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
            for t in threads:
                t.start()
            for t in threads:
                t.join(timeout=120)
    finally:
        for d in devs:
            d.close()
    assert not errors, errors
    return results


@pytest.mark.parametrize("persistent", [True, False])
@pytest.mark.parametrize("world", [2, 3])
def test_peer_exchange_ice_on_one_gpu(world, persistent, gm12878):
    m = _ice_input(gm12878)
    with DeviceCSR.from_scipy(m) as dev:
        bias1, it1, dev1 = dev.ice(500, 1e-5)
    for bias, iters, devn in _run_peer(m, world, lambda d: d.ice(500, 1e-5), persistent):
        assert int(iters[0]) == int(it1[0])
        np.testing.assert_array_equal(bias, bias1)
        assert devn[0] == dev1[0]


@pytest.mark.parametrize("persistent", [True, False])
@pytest.mark.parametrize("world", [2, 3])
def test_peer_exchange_kr_on_one_gpu(world, persistent, gm12878):
    m = full_from_pixels(gm12878)
    with DeviceCSR.from_scipy(m) as dev:
        x1, outer1, mv1, res1 = dev.kr()
    for x, outer, mv, res in _run_peer(m, world, lambda d: d.kr(), persistent):
        assert (outer, mv) == (outer1, mv1)
        np.testing.assert_array_equal(x, x1)
        assert res == res1


def test_peer_exchange_ice_then_kr_on_the_same_handles(gm12878):
    """the exchange sequence number carries over from one solver launch to the next (early-stopped loops included)"""
    m = _ice_input(gm12878)
    with DeviceCSR.from_scipy(m) as dev:
        bias1, it1, _ = dev.ice(500, 1e-5)
        x1, _o, mv1, _r = dev.kr()
        bias2, it2, _ = dev.ice(7, 0.0)

    def both(d):
        b, it, _ = d.ice(500, 1e-5)
        x, _o, mv, _r = d.kr()
        b2, it2, _ = d.ice(7, 0.0)
        return b, int(it[0]), x, mv, b2, int(it2[0])

    for b, it, x, mv, b2, i2 in _run_peer(m, 2, both):
        assert it == int(it1[0]) and mv == mv1 and i2 == int(it2[0]) == 7
        np.testing.assert_array_equal(b, bias1)
        np.testing.assert_array_equal(x, x1)
        np.testing.assert_array_equal(b2, bias2)
