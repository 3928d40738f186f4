"""hb_csr_pack_values: the optional narrow (uint16 / float) copy of the values that the streaming kernels read must
be LOSSLESS -- every result bit-identical to the fp64 layout -- and must refuse inputs it cannot represent exactly."""
import numpy as np
import pytest

from hicexplorer_b200.store import DeviceCSR
from tests.helpers import random_hic

pytestmark = pytest.mark.gpu


def _run_all(dev):
    bias, iters, devn = dev.ice(150, 1e-7)
    values = dev.ice_scaled_values()
    x, outer, mv, res = dev.kr()
    return bias, int(iters[0]), values, x, mv


@pytest.mark.parametrize("scale,expect_kind", [(1.0, 1), (0.5, 2), (300.0, 2), (1.0 / 3.0, 0)])
def test_packed_layouts_are_bit_identical(scale, expect_kind):
    m = random_hic(900, 0.08, 5)
    m.data *= scale                      # counts | halves (exact in fp32) | > 65535 | thirds (not representable)
    with DeviceCSR.from_scipy(m) as dev:
        ref = _run_all(dev)
        kind = dev.pack_values()
        assert kind == expect_kind
        got = _run_all(dev)
    for a, b in zip(ref, got):
        if isinstance(a, np.ndarray):
            np.testing.assert_array_equal(a, b)
        else:
            assert a == b


def test_pack_is_dropped_when_values_change():
    m = random_hic(300, 0.1, 9)
    mask = np.zeros(300, dtype=np.uint8)
    mask[::7] = 1
    with DeviceCSR.from_scipy(m) as dev:
        assert dev.pack_values() == 1
        dev.compact(mask)                # new value array: the stale copy must not be used
        bias, iters, _ = dev.ice(50, 1e-5)
    keep = np.flatnonzero(mask == 0)
    with DeviceCSR.from_scipy(m[keep][:, keep].tocsr()) as dev2:
        bias2, iters2, _ = dev2.ice(50, 1e-5)
    np.testing.assert_array_equal(bias, bias2)


def test_nan_and_negative_values_do_not_pack():
    m = random_hic(200, 0.1, 2)
    m.data[3] = -1.0
    with DeviceCSR.from_scipy(m) as dev:
        assert dev.pack_values() == 2    # negative integers: not uint16, but exact in fp32
    m.data[5] = np.nan
    with DeviceCSR.from_scipy(m) as dev:
        assert dev.pack_values() == 0
