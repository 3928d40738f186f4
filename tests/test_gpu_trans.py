"""--transCutoff / --inflationCutoff (hicCorrectMatrix.py:695-699, 765-775) on the device."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200.pipeline import ice_pipeline
from hicexplorer_b200.store import DeviceCSR
from oracle import ice_oracle
from tests.helpers import random_hic

pytestmark = pytest.mark.gpu


def _multi_chrom(n, seed, nchr=4, integer=True, trans_density=0.02):
    """random symmetric matrix with `nchr` chromosomes and scattered trans entries (a few of them huge)"""
    rng = np.random.default_rng(seed)
    m = random_hic(n, 0.15, seed, integer=integer)
    k = int(n * n * trans_density)
    i, j = rng.integers(0, n, k), rng.integers(0, n, k)
    v = np.floor(rng.exponential(3.0, k)) + 1
    v[rng.random(k) < 0.01] *= 500
    if not integer:
        v = v * rng.random(k) - 0.5            # negative values too
    t = sparse.coo_matrix((v, (i, j)), shape=(n, n)).tocsr()
    t = sparse.triu(t, 1)
    m = sparse.csr_matrix(m + t + t.T)
    m.sort_indices()
    offs = np.concatenate([[0], np.sort(rng.choice(np.arange(1, n), nchr - 1, replace=False)), [n]]).astype(np.int64)
    return m, offs


def _trans_values(m, offs):
    coo = m.tocoo()
    trans = np.searchsorted(offs, coo.row, side="right") != np.searchsorted(offs, coo.col, side="right")
    return coo.data[trans]


@pytest.mark.parametrize("n,seed,integer", [(300, 0, True), (1500, 1, True), (800, 2, False), (64, 3, True)])
@pytest.mark.parametrize("q", [99.9995, 50.0, 0.0, 100.0, 37.3])
def test_trans_percentile_is_numpys(n, seed, integer, q):
    m, offs = _multi_chrom(n, seed, integer=integer)
    want = np.percentile(_trans_values(m, offs), q)
    with DeviceCSR.from_scipy(m) as dev:
        got = dev.trans_percentile(offs, q)
    assert got == want                          # bit-exact: exact order statistics + numpy's interpolation


def test_trans_percentile_without_trans_entries():
    m = random_hic(200, 0.2, 4)
    with DeviceCSR.from_scipy(m) as dev:
        assert dev.trans_percentile(np.array([0, 200]), 99.0) is None


def test_trans_clip_matches_oracle():
    m, offs = _multi_chrom(1000, 5)
    want, max_inter = ice_oracle.trunc_trans(m, offs, 1.0, clip=True)
    want.sort_indices()
    with DeviceCSR.from_scipy(m) as dev:
        got_max = dev.trans_percentile(offs, 100 - 1.0)
        clipped = dev.trans_clip(offs, got_max)
        got = dev.download()
    assert got_max == max_inter
    assert clipped == int(np.count_nonzero(_trans_values(m, offs) >= max_inter))
    assert np.array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)
    assert abs(got - got.T).max() == 0


@pytest.mark.parametrize("clip", [False, True])
def test_pipeline_with_trans_and_inflation_cutoff(clip):
    m, offs = _multi_chrom(1200, 6, nchr=3)
    kw = dict(chr_offsets=offs, trans_cutoff=5.0, clip_trans=clip, inflation_cutoff=1.6)
    ref = ice_oracle.correct_ice_pipeline(m, -2.0, 4.0, max_iter=300, **kw)
    got = ice_pipeline(m, (-2.0, 4.0), max_iter=300, file_layout=True, **kw)
    assert got["max_inter"] == ref["max_inter"]
    assert 0 < len(ref["inflated_rel"]) < 600           # the case exercises the mask
    assert np.array_equal(got["inflated_rel"], ref["inflated_rel"])
    assert np.array_equal(got["kept"], ref["kept"])
    np.testing.assert_allclose(got["bias"], ref["ice"]["bias"], rtol=1e-9)
    up = got["upper"]
    assert np.array_equal(up.indptr, ref["upper"].indptr) and np.array_equal(up.indices, ref["upper"].indices)
    np.testing.assert_allclose(up.data, ref["upper"].data, rtol=1e-8)


def test_scaled_row_sums_are_the_corrected_matrix_row_sums():
    m = random_hic(900, 0.1, 7)
    with DeviceCSR.from_scipy(m) as dev:
        dev.ice(200, 1e-5)
        sums = dev.ice_scaled_row_sums()
        values = dev.ice_scaled_values()
    w = sparse.csr_matrix((values, m.indices, m.indptr), shape=m.shape)
    np.testing.assert_allclose(sums, np.asarray(w.sum(axis=1)).ravel(), rtol=1e-12)


def test_inflation_needs_trans_cutoff():
    m, offs = _multi_chrom(300, 8)
    with pytest.raises(ValueError, match="trans_cutoff"):
        ice_pipeline(m, (-2.0, 4.0), chr_offsets=offs, inflation_cutoff=2.0, file_layout=True)
