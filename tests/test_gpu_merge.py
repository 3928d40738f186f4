"""Bin merging (hicMergeMatrixBins --numBins k) on the GPU vs the reference's golden file and the CPU oracle."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200 import hdf5lite
from hicexplorer_b200.hicmatrix_lite import hiCMatrix
from oracle import merge_oracle
from tests.conftest import load_golden
from tests.helpers import random_hic

pytestmark = pytest.mark.gpu


def _input(tmp_path):
    z = load_golden("merge_small.npz")
    n = len(z["in_indptr"]) - 1
    ma = hiCMatrix()
    up = sparse.csr_matrix((z["in_data"], z["in_indices"], z["in_indptr"]), shape=(n, n))
    ma.matrix = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    ma.cut_intervals = [(c.decode(), int(s), int(e), float(x)) for c, s, e, x in
                        zip(z["in_chr"], z["in_start"], z["in_end"], z["in_extra"])]
    ma._index_intervals()
    path = str(tmp_path / "small_test_matrix.h5")
    ma.save(path)
    return z, path


def test_merge_matrix_bins_matches_reference_golden(tmp_path):
    """mirrors test_correct_matrix in test/general/test_hicMergeMatrixBins.py:15-27 (exact data + intervals)"""
    from hicexplorer_b200.hicMergeMatrixBins import main
    z, src = _input(tmp_path)
    out = str(tmp_path / "merged.h5")
    main("--matrix {} --numBins 5 --outFileName {}".format(src, out).split())
    f = hdf5lite.File(out)
    np.testing.assert_array_equal(f["matrix/data"].read(), z["out_data"])          # bit-exact (integer counts)
    np.testing.assert_array_equal(f["matrix/indices"].read(), z["out_indices"])
    np.testing.assert_array_equal(f["matrix/indptr"].read(), z["out_indptr"])
    assert [c.decode() for c in f["intervals/chr_list"].read()] == [c.decode() for c in z["out_chr"]]
    np.testing.assert_array_equal(f["intervals/start_list"].read(), z["out_start"])
    np.testing.assert_array_equal(f["intervals/end_list"].read(), z["out_end"])
    np.testing.assert_allclose(f["intervals/extra_list"].read(), z["out_extra"], rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(f["nan_bins"].read(), z["out_nan_bins"])


@pytest.mark.parametrize("n,k,seed", [(500, 2, 0), (1203, 7, 1), (64, 32, 2), (777, 3, 3), (2000, 100, 4), (1500, 33, 5)])
def test_merge_random_matrices_match_oracle(n, k, seed):
    from hicexplorer_b200.store import DeviceCSR
    m = random_hic(n, 0.1, seed, empty_frac=0.02)
    sizes = [n // 3, n // 3, n - 2 * (n // 3)]
    iv = []
    for c, L in enumerate(sizes):
        iv += [("chr%d" % c, i * 100, (i + 1) * 100, 1.0) for i in range(L)]
    ref, new_iv, _nan, bin_map = merge_oracle.merge_bins(m, iv, k)
    with DeviceCSR.from_scipy(m) as dev:
        dev.merge_bins(bin_map, len(new_iv))
        got = dev.download()
        ratio = dev.symmetry_ratio()
        again = dev.download()
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    np.testing.assert_array_equal(got.data, ref.data)
    np.testing.assert_array_equal(again.data, ref.data)
    assert ratio == 0.0


def test_merge_then_correct_workflow():
    """merge -> ICE on the same device store (the reference's documented order: merge first, then correct)"""
    from hicexplorer_b200.store import DeviceCSR
    from oracle import ice_oracle
    m = random_hic(900, 0.1, 11)
    iv = [("chr1", i * 10, (i + 1) * 10, 1.0) for i in range(900)]
    ref_m, new_iv, _nan, bin_map = merge_oracle.merge_bins(m, iv, 4)
    ref = ice_oracle.ice(ref_m.astype(np.float64), max_iter=100)
    with DeviceCSR.from_scipy(m) as dev:
        dev.merge_bins(bin_map, len(new_iv))
        bias, iters, _ = dev.ice(100, 1e-5)
    np.testing.assert_allclose(bias, ref["bias"], rtol=1e-9)
    assert abs(int(iters[0]) - ref["iterations"]) <= 1


@pytest.mark.parametrize("n,k,seed", [(200, 3, 0), (777, 5, 1), (64, 7, 2), (30, 1, 3)])
def test_running_window_matches_oracle(n, k, seed):
    """--runningWindow (hicMergeMatrixBins.py:88-186): sum of the k^2 shifted copies of the upper triangle on the device"""
    from hicexplorer_b200.hicMergeMatrixBins import running_window_merge
    m = random_hic(n, 0.1, seed, empty_frac=0.03).astype(np.int32)
    want = merge_oracle.running_window(m, k)
    hic = hiCMatrix()
    hic.matrix = m.copy()
    hic.cut_intervals = [("chr1", i * 10, (i + 1) * 10, 1.0) for i in range(n)]
    hic._index_intervals()
    got = running_window_merge(hic, k).matrix
    got.sort_indices()
    assert got.dtype == want.dtype
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)


def test_shifted_copy_and_symmetrize_in_place():
    from hicexplorer_b200.store import DeviceCSR
    m = random_hic(300, 0.1, 9)
    with DeviceCSR.from_scipy(m) as dev:
        for dr, dc in [(0, 0), (1, -2), (-3, 3), (5, 0)]:
            with dev.shifted_copy(dr, dc) as sh:
                got = sh.download()
            coo = m.tocoo()
            r, c = coo.row + dr, coo.col + dc
            ok = (r >= 0) & (r < 300) & (c >= 0) & (c < 300)
            want = sparse.csr_matrix((coo.data[ok], (r[ok], c[ok])), shape=m.shape)
            want.sort_indices()
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            np.testing.assert_array_equal(got.data, want.data)
        dev.keep_upper()
        dev.symmetrize_upper()
        back = dev.download()
    assert np.array_equal(back.indices, m.indices)
    np.testing.assert_array_equal(back.data, m.data)
