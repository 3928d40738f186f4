"""csrc/hb_hostpack.cpp on the CPU: the host side of the narrow ingest converts fp64 counts to uint16 only when the
round trip is bit-exact (compiled here with g++ into a scratch library; the product links the same file into
libhicbalance.so)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pack(tmp_path_factory):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("hostpack") / "libhostpack.so")
    subprocess.check_call([cxx, "-O3", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "hicexplorer_b200", "csrc", "hb_hostpack.cpp")])
    lib = ctypes.CDLL(so)
    lib.hb_host_pack_u16.restype = ctypes.c_int
    lib.hb_host_pack_u16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]

    def run(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        out = np.full(a.size, 0xABCD, dtype=np.uint16)
        ok = lib.hb_host_pack_u16(a.ctypes.data, out.ctypes.data, a.size)
        return bool(ok), out
    return run


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 31, 33, 1000, 4099])
def test_counts_round_trip_exactly(pack, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 65536, n).astype(np.float64)
    if n:
        a[0], a[-1] = 65535.0, 0.0
    ok, out = pack(a)
    assert ok and np.array_equal(out.astype(np.float64), a)


@pytest.mark.parametrize("bad", [65536.0, -1.0, 0.5, 1e-300, -0.0, np.nan, np.inf, -np.inf, 2.0 ** 31, -2.0 ** 31, 2.0 ** 40,
                                 65535.000000000007])
@pytest.mark.parametrize("n,pos", [(1, 0), (16, 5), (16, 15), (40, 0), (40, 31), (40, 32), (40, 39), (1000, 512)])
def test_anything_else_is_refused(pack, bad, n, pos):
    a = np.arange(n, dtype=np.float64) % 1000
    a[pos] = bad
    ok, _ = pack(a)
    assert not ok


def test_unaligned_input_and_output(pack):
    base = np.arange(1001, dtype=np.float64) % 500
    ok, out = pack(base[1:])                       # 8-byte aligned only
    assert ok and np.array_equal(out, (base[1:]).astype(np.uint16))
