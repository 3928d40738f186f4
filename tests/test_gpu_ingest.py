"""hb_csr_create's narrow ingest (csrc/hb_xfer.cu: hb_ingest_values + csrc/hb_hostpack.cpp): fp64 counts are converted
to uint16 by host threads, 2 bytes per entry cross PCIe, the device writes the fp64 store and the uint16 stream from
them.  Lossless or refused: the store must equal the caller's matrix bit for bit either way, and the solvers must give
the same bits as on a store that never saw the narrow path."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200 import _lib
from hicexplorer_b200.store import DeviceCSR
from tests.helpers import random_hic

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _narrow_on(monkeypatch):
    """without the variable the library decides by the number of host threads this rank has (>= 8: on)"""
    monkeypatch.setenv("HB_NARROW_INGEST", "1")


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _info(dev):
    kind, nbytes = ctypes.c_int(-1), ctypes.c_int64(-1)
    _lib.check(_lib.lib().hb_csr_ingest_info(dev.handle, ctypes.byref(kind), ctypes.byref(nbytes)))
    return kind.value, nbytes.value


def _same_bits(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float64).view(np.uint64), np.asarray(b, dtype=np.float64).view(np.uint64))


def test_counts_arrive_narrow_and_the_store_equals_the_input():
    m = random_hic(3000, 0.05, 5, integer=True)
    m.data[::7] = 65535.0
    m.data[3::11] = 0.0          # explicit zeros are counts too
    with DeviceCSR.from_scipy(m) as dev:
        kind, nbytes = _info(dev)
        assert kind == 1
        assert nbytes == 8 * (m.shape[0] + 1) + 4 * m.nnz + 2 * m.nnz
        got = dev.download()
        assert dev.pack_values() == 1            # nothing to do: the copy exists already
        bias, iters, _ = dev.ice(40, 1e-6)
    assert np.array_equal(got.indptr, m.indptr) and np.array_equal(got.indices, m.indices) and _same_bits(got.data, m.data)
    # the same matrix through the int32 entry point never sees the narrow path: fp64 kernels, same bits
    mi = sparse.csr_matrix((m.data.astype(np.int32), m.indices, m.indptr), shape=m.shape)
    with DeviceCSR.from_scipy(mi) as dev:
        assert _info(dev)[0] == 0
        bias2, iters2, _ = dev.ice(40, 1e-6)
    assert int(iters[0]) == int(iters2[0]) and _same_bits(bias, bias2)


@pytest.mark.parametrize("bad", [65536.0, 0.5, -1.0, -0.0, np.nan, np.inf, 2.0 ** 31])
@pytest.mark.parametrize("where", ["head", "middle", "tail"])
def test_anything_that_is_not_a_small_count_takes_the_fp64_copy(bad, where):
    m = random_hic(1500, 0.05, 9, integer=True)
    pos = {"head": 3, "middle": m.nnz // 2, "tail": m.nnz - 2}[where]
    m.data[pos] = bad
    with DeviceCSR.from_scipy(m) as dev:
        kind, nbytes = _info(dev)
        got = dev.download()
    assert kind == 0 and nbytes == 8 * (m.shape[0] + 1) + 12 * m.nnz
    assert _same_bits(got.data, m.data)


def test_row_blocks_arrive_narrow_too():
    m = random_hic(2000, 0.08, 4, integer=True)
    lo, hi = 613, 1401                       # a block whose first entry is not 4-aligned in the global entry order
    with DeviceCSR.from_scipy(m, row_range=(lo, hi)) as dev:
        assert _info(dev)[0] == 1
        got = dev.download()
    want = m[lo:hi]
    assert np.array_equal(got.indices, want.indices) and _same_bits(got.data, want.data)


SCRIPT = r"""
import ctypes, numpy as np, sys
sys.path.insert(0, %r)
from hicexplorer_b200 import _lib
from hicexplorer_b200.store import DeviceCSR
from tests.helpers import random_hic
m = random_hic(2500, 0.06, 2, integer=True)          # ~ 375 000 entries = hundreds of 1000-entry chunks on 3 workers
def info(dev):
    k, b = ctypes.c_int(-1), ctypes.c_int64(-1)
    _lib.check(_lib.lib().hb_csr_ingest_info(dev.handle, ctypes.byref(k), ctypes.byref(b)))
    return k.value
with DeviceCSR.from_scipy(m) as dev:
    assert info(dev) == 1
    got = dev.download()
    bias, iters, _ = dev.ice(30, 1e-6)
assert np.array_equal(got.data.view(np.uint64), m.data.view(np.uint64)) and np.array_equal(got.indices, m.indices)
for pos in (700, m.nnz // 3, m.nnz - 700):              # refused by a worker in the middle of the transfer
    b = m.copy(); b.data[pos] = 0.25
    with DeviceCSR.from_scipy(b) as dev:
        assert info(dev) == 0
        got = dev.download()
    assert np.array_equal(got.data.view(np.uint64), b.data.view(np.uint64))
print("OK", int(iters[0]), float(bias.sum()))
""" % ROOT


def test_many_small_chunks_and_a_refusal_in_the_middle():
    env = dict(os.environ, HB_INGEST_CHUNK="1000", HB_INGEST_WORKERS="3", HB_NARROW_INGEST="1")
    out = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
    env = dict(os.environ, HB_NARROW_INGEST="0")
    ref = subprocess.run([sys.executable, "-c", SCRIPT.replace("assert info(dev) == 1", "assert info(dev) == 0")], env=env,
                         capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0 and ref.stdout == out.stdout, ref.stdout + ref.stderr     # same iterations, same bias sum


FEEDER_SCRIPT = r"""
import ctypes, numpy as np, sys, torch
sys.path.insert(0, %r)
from scipy import sparse
from hicexplorer_b200 import _lib
from hicexplorer_b200.store import DeviceCSR
from tests.helpers import random_hic
m = random_hic(2500, 0.06, 2, integer=True)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t
keep = []
def pinned_csr(m):
    d, i, p = pinned(m.data), pinned(m.indices), pinned(m.indptr)
    keep.extend([d, i, p])
    return sparse.csr_matrix((d.numpy(), i.numpy(), p.numpy()), shape=m.shape)
def info(dev):
    k, b = ctypes.c_int(-1), ctypes.c_int64(-1)
    _lib.check(_lib.lib().hb_csr_ingest_info(dev.handle, ctypes.byref(k), ctypes.byref(b)))
    return k.value, b.value
mp = pinned_csr(m)
fixed = 8 * (m.shape[0] + 1) + 4 * m.nnz
shares = []
for rep in range(3):
    with DeviceCSR.from_scipy(mp) as dev:
        kind, nbytes = info(dev)
        got = dev.download()
        bias, iters, _ = dev.ice(30, 1e-6)
    assert kind == 1
    assert fixed + 2 * m.nnz <= nbytes <= fixed + 8 * m.nnz and (nbytes - fixed - 2 * m.nnz) %% 6 == 0
    shares.append((nbytes - fixed - 2 * m.nnz) // 6)        # entries that went as fp64
    assert np.array_equal(got.data.view(np.uint64), m.data.view(np.uint64)) and np.array_equal(got.indices, m.indices)
for pos in (700, m.nnz // 3, m.nnz - 5000, m.nnz - 700):    # refused by a worker (front) or by the device pass (tail)
    b = m.copy(); b.data[pos] = 0.25
    with DeviceCSR.from_scipy(pinned_csr(b)) as dev:
        kind, nbytes = info(dev)
        got = dev.download()
    assert kind == 0 and fixed + 2 * m.nnz <= nbytes <= fixed + 8 * m.nnz
    assert np.array_equal(got.data.view(np.uint64), b.data.view(np.uint64))
print("OK", int(iters[0]), float(bias.sum()), "direct entries per run:", shares)
""" % ROOT


def test_direct_feeder_shares_the_upload_with_the_narrow_side():
    """pinned source, one slow converter thread, small chunks: the feeder takes a share of the chunks from the back and
    sends them as fp64; the store and the uint16 copy must come out the same, refusals included"""
    env = dict(os.environ, HB_INGEST_CHUNK="1000", HB_INGEST_WORKERS="1", HB_NARROW_INGEST="1")
    out = subprocess.run([sys.executable, "-c", FEEDER_SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
    ref = subprocess.run([sys.executable, "-c", SCRIPT], env=dict(os.environ, HB_NARROW_INGEST="1", HB_INGEST_CHUNK="1000"),
                         capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0 and ref.stdout.split()[1:3] == out.stdout.split()[1:3], ref.stdout + ref.stderr + out.stdout
