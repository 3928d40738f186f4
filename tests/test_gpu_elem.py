"""hicNormalize / hicSumMatrices on the device against the reference's golden files and the oracle."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200 import hdf5lite
from hicexplorer_b200.hicmatrix_lite import hiCMatrix
from hicexplorer_b200.store import DeviceCSR
from oracle import normalize_oracle
from tests.conftest import load_golden
from tests.helpers import random_hic

pytestmark = pytest.mark.gpu


def _files(tmp_path):
    z = load_golden("normalize_small.npz")
    n = int(z["n"])
    iv = [(c.decode(), int(s), int(e), float(x)) for c, s, e, x in zip(z["chr"], z["start"], z["end"], z["extra"])]
    paths = []
    for key, ext in [("one", ".h5"), ("two", ".h5"), ("one", ".cool")]:
        up = sparse.csr_matrix((z[key + "_data"], z[key + "_indices"], z[key + "_indptr"]), shape=(n, n))
        ma = hiCMatrix()
        ma.matrix = sparse.csr_matrix(up + sparse.triu(up, 1).T)
        ma.cut_intervals = iv
        ma._index_intervals()
        p = str(tmp_path / ("in_" + key + ext))
        ma.save(p)
        paths.append(p)
    return z, paths


def _h5(path):
    f = hdf5lite.File(path)
    return f["matrix/data"].read(), f["matrix/indices"].read(), f["matrix/indptr"].read()


@pytest.mark.parametrize("mode,keys", [("smallest", ("smallest_one", "smallest_two")),
                                       ("norm_range", ("norm_range_one", "norm_range_two"))])
def test_hicNormalize_matches_reference_golden(tmp_path, mode, keys):
    """mirrors test_normalize_smallest / test_normalize_norm_range (test_hicNormalize.py:23-131): exact data"""
    from hicexplorer_b200.hicNormalize import main
    z, (one, two, _cool) = _files(tmp_path)
    o1, o2 = str(tmp_path / "o1.h5"), str(tmp_path / "o2.h5")
    main("--matrices {} {} --normalize {} -o {} {}".format(one, two, mode, o1, o2).split())
    for out, key in zip((o1, o2), keys):
        d, i, p = _h5(out)
        assert d.dtype == np.float32
        assert np.array_equal(i, z[key + "_indices"]) and np.array_equal(p, z[key + "_indptr"])
        np.testing.assert_array_equal(d, z[key + "_data"])


def test_hicNormalize_multiplicative_cool(tmp_path):
    """mirrors test_normalize_multiplicative_h5_cool (:239-258)"""
    from hicexplorer_b200.hicNormalize import main
    z, (_one, _two, cool) = _files(tmp_path)
    out = str(tmp_path / "by2.cool")
    main("--matrices {} --normalize multiplicative --multiplicativeValue 2 -o {}".format(cool, out).split())
    up = hiCMatrix(out).upper_for_device()
    assert np.array_equal(up.indices, z["by2_indices"]) and np.array_equal(up.indptr, z["by2_indptr"])
    np.testing.assert_array_equal(up.data, z["by2_data"])


@pytest.mark.parametrize("mode", ["smallest", "norm_range", "multiplicative"])
def test_normalize_random_float_matrices_match_oracle(mode):
    from hicexplorer_b200.hicNormalize import normalize_matrices
    mats = [random_hic(700, 0.1, s, integer=False) for s in (1, 2, 3)]
    if mode != "smallest":                    # (a NaN would make every sum ratio of `smallest` NaN: degenerate case)
        mats[1].data[[5, 77]] = np.nan        # scrubbed to 0 and eliminated (asymmetric on purpose: element-wise path)
        mats[2].data[11] = np.inf
    ref = normalize_oracle.normalize(mats, mode, multiplicative_value=0.37, threshold=0.02)
    hics = []
    for m in mats:
        h = hiCMatrix()
        h.matrix = m.copy()
        hics.append(h)
    got = normalize_matrices(hics, mode, multiplicative_value=0.37, threshold=0.02)
    for g, r in zip(got, ref):
        r.sort_indices()
        assert g.matrix.data.dtype == np.float32
        assert np.array_equal(g.matrix.indptr, r.indptr) and np.array_equal(g.matrix.indices, r.indices)
        np.testing.assert_array_equal(g.matrix.data, r.data)         # bit-exact float32


def test_value_stats():
    m = random_hic(500, 0.2, 4, integer=False)
    with DeviceCSR.from_scipy(m) as dev:
        s, mn, mx = dev.value_stats()
    assert abs(s - m.sum()) <= 1e-9 * abs(m.sum())
    assert mn == np.float32(m.data.min()) and mx == np.float32(m.data.max())


@pytest.mark.parametrize("n,seed", [(300, 0), (1500, 1), (40, 2)])
def test_add_matches_scipy(n, seed):
    a = random_hic(n, 0.1, seed)
    b = random_hic(n, 0.05, seed + 10, empty_frac=0.05)
    c = random_hic(n, 0.3, seed + 20)
    b.data[::7] *= -1
    a2 = sparse.csr_matrix(a - b)             # exact cancellations: a - b + b has cells that sum to 0
    a2.sort_indices()
    want = normalize_oracle.sum_matrices([a2, b, c])
    want.sort_indices()
    with DeviceCSR.from_scipy(a2) as total, DeviceCSR.from_scipy(b) as db, DeviceCSR.from_scipy(c) as dc:
        total.add(db)
        total.add(dc)
        got = total.download()
        assert db.download().nnz == b.nnz
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)


def test_hicSumMatrices_cli(tmp_path):
    from hicexplorer_b200.hicSumMatrices import main
    z, (one, two, cool) = _files(tmp_path)
    out = str(tmp_path / "sum.h5")
    main("--matrices {} {} {} --outFileName {}".format(one, two, cool, out).split())
    a, b = hiCMatrix(one).matrix, hiCMatrix(two).matrix
    want = sparse.triu(a + b + a, 0, format="csr")
    want.sort_indices()
    d, i, p = _h5(out)
    assert np.array_equal(i, want.indices) and np.array_equal(p, want.indptr)
    np.testing.assert_array_equal(d, want.data)
    with pytest.raises(SystemExit):
        bad = hiCMatrix(one)
        bad.reorderChromosomes(list(reversed(bad.getChrNames())))
        badp = str(tmp_path / "bad.h5")
        bad.save(badp)
        main("--matrices {} {} --outFileName {}".format(one, badp, out).split())


def test_edge_cases_empty_and_tiny_stores():
    """empty matrix, 1 x 1 matrix, rows that vanish completely: the element-wise entry points must not trip"""
    from hicexplorer_b200.hicTransform import obs_exp_transform
    empty = sparse.csr_matrix((50, 50), dtype=np.float64)
    with DeviceCSR.from_scipy(empty) as dev, DeviceCSR.from_scipy(empty) as other:
        s, mn, mx = dev.value_stats(scrub=True)
        assert s == 0.0 and np.isnan(mn) and np.isnan(mx)
        dev.normalize("multiply", a=3.0)
        dev.eliminate_zeros()
        dev.add(other)
        sums, counts = dev.distance_stats()
        assert dev.nnz == 0 and sums.sum() == 0 and counts.sum() == 0
    one = sparse.csr_matrix(np.array([[4.0]]))
    with DeviceCSR.from_scipy(one) as dev, DeviceCSR.from_scipy(one) as other:
        dev.add(other)
        assert dev.download()[0, 0] == 8.0
        dev.normalize("divide", a=2.0)
        assert dev.download()[0, 0] == 4.0
        assert dev.value_stats() == (4.0, 4.0, 4.0)
    got = obs_exp_transform(one.astype(np.int32), "obs_exp", np.array([0, 1]))
    assert got.nnz == 1 and got[0, 0] == 2          # expected = 4 / (n + 1 - 0) = 2 -> 4 / 2
    m = random_hic(300, 0.1, 31)
    neg = sparse.csr_matrix(-m)
    with DeviceCSR.from_scipy(m) as dev, DeviceCSR.from_scipy(neg) as other:
        dev.add(other)                               # everything cancels: scipy would return an empty matrix too
        assert dev.nnz == 0 and (m + neg).nnz == 0
        assert dev.download().shape == (300, 300)
    with DeviceCSR.from_scipy(m) as dev:
        dev.normalize("multiply", a=1.0, threshold=1e9)        # every value below the threshold
        dev.eliminate_zeros()
        assert dev.nnz == 0
        bias, iters, devn = dev.ice(3, 1e-5)                   # and the store is still usable
        assert np.all(bias == 0.0) or np.all(np.isnan(bias)) or bias.shape == (300,)


@pytest.mark.parametrize("n,band,per_row,far", [(150000, 3000, 5, 30000), (70000, 66000, 3, 0), (3000, 3000, 400, 0)])
def test_add_on_wide_rows(n, band, per_row, far):
    """rows whose columns span many accumulation windows, partial overlap of the two patterns, cells that cancel,
    explicit zeros -- against scipy"""
    rng = np.random.default_rng(n)

    def sym(seed):
        g = np.random.default_rng(seed)
        r = np.repeat(np.arange(n), per_row)
        c = np.minimum(r + g.integers(0, band, size=len(r)), n - 1)
        fr, fc = g.integers(0, n, size=far), g.integers(0, n, size=far)
        r, c = np.concatenate([r, np.minimum(fr, fc)]), np.concatenate([c, np.maximum(fr, fc)])
        v = g.integers(1, 50, size=len(r)).astype(np.float64)
        up = sparse.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        up.sum_duplicates()
        m = sparse.csr_matrix(up + sparse.triu(up, 1).T)
        m.sort_indices()
        return m
    a, b = sym(1), sym(2)
    common = a.multiply(b > 0).tocsr()                       # cells both matrices store (a's values there)
    common.data[common.data > 10] = 0.0                      # (never compare a sparse matrix with <=: the result is dense)
    common.eliminate_zeros()
    bc = b.multiply(common > 0).tocsr()                      # b's values at the chosen cells
    b = sparse.csr_matrix(b - bc - common)                   # there b = -a now: a + b cancels exactly; pattern of b kept
    b.sort_indices()
    a.data[rng.integers(0, a.nnz, size=50)] = 0.0            # explicit zeros that meet nothing, or something
    want = sparse.csr_matrix(a + b)
    want.sort_indices()
    with DeviceCSR.from_scipy(a) as da, DeviceCSR.from_scipy(b) as db:
        da.add(db)
        got = da.download()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)
