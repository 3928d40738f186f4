"""The device generator of benchmark inputs vs its numpy restatement (bit-exact), plus structural checks."""
import numpy as np
import pytest

from oracle import ice_oracle
from oracle.synth_ref import synth_matrix

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chrom_bins,kw", [
    ([300, 200, 57], dict(seed=0, band=40, d0=6.0, n_trans=5)),
    ([1000], dict(seed=3, band=1000, d0=20.0, n_trans=0, empty_frac=0.05)),
    ([64, 64, 64, 64], dict(seed=9, band=10, d0=3.0, n_trans=32, low_frac=0.2)),
])
def test_generator_is_bit_identical_to_numpy_restatement(chrom_bins, kw):
    from hicexplorer_b200.store import DeviceCSR
    ref = synth_matrix(chrom_bins, **kw)
    with DeviceCSR.synthetic(chrom_bins, **kw) as dev:
        got = dev.download()
        assert dev.symmetry_ratio() == 0.0
    assert np.array_equal(got.indptr, ref.indptr)
    assert np.array_equal(got.indices, ref.indices)
    assert np.array_equal(got.data, ref.data)
    assert np.all(got.data == np.floor(got.data)) and got.data.min() >= 1


def test_row_block_generation_matches_full():
    from hicexplorer_b200.store import DeviceCSR
    chrom_bins = [500, 300]
    kw = dict(seed=1, band=60, d0=8.0, n_trans=4)
    with DeviceCSR.synthetic(chrom_bins, **kw) as dev:
        full = dev.download()
    for a, b in ((0, 123), (123, 640), (640, 800)):
        with DeviceCSR.synthetic(chrom_bins, row_range=(a, b), **kw) as dev:
            blk = dev.download()
        ref = full[a:b]
        assert np.array_equal(blk.indices, ref.indices) and np.array_equal(blk.data, ref.data)


def test_filter_and_ice_on_synthetic_matrix_match_oracle():
    from hicexplorer_b200.pipeline import ice_pipeline
    chrom_bins = [900, 700, 400]
    m = synth_matrix(chrom_bins, seed=5, band=200, d0=25.0, n_trans=6)
    ref = ice_oracle.correct_ice_pipeline(m, -2.0, 5.0, max_iter=300)
    got = ice_pipeline(m, (-2.0, 5.0), max_iter=300)
    assert np.array_equal(got["zero_bins"], ref["zero_bins"]) and len(ref["zero_bins"]) > 0
    assert np.array_equal(got["outliers_rel"], ref["outliers_rel"]) and len(ref["outliers_rel"]) > 0
    assert abs(int(got["iterations"][0]) - ref["ice"]["iterations"]) <= 1
    np.testing.assert_allclose(got["bias"], ref["ice"]["bias"], rtol=1e-9)
