"""hicTransform obs/exp on the device against golden vectors produced by the reference's own
functions (tests/golden/make_golden_transform.py), the reference's golden files, and the CPU oracle."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200.hicmatrix_lite import hiCMatrix
from hicexplorer_b200.hicTransform import correlation_transform, main, obs_exp_transform
from oracle import transform_oracle
from tests.conftest import load_golden
from tests.helpers import random_hic
from tests.test_oracle import TRANSFORM_CASES, _transform_inputs, _upper

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,which,method,pc,lig", TRANSFORM_CASES)
def test_obs_exp_equals_reference_functions(key, which, method, pc, lig):
    t, ma, mb = _transform_inputs()
    m, offs = (ma, t["offsets"]) if which == "a" else (mb, t["b_offsets"])
    got = _upper(obs_exp_transform(m, method, offs, per_chromosome=pc, ligation_factor=lig))
    assert np.array_equal(got.indptr, t[key + "_indptr"]) and np.array_equal(got.indices, t[key + "_indices"])
    np.testing.assert_array_equal(got.data, t[key + "_data"])            # integer counts: bit-exact


@pytest.mark.parametrize("method,pc,lig", [("obs_exp", False, False), ("obs_exp", True, False), ("obs_exp_non_zero", False, True),
                                           ("obs_exp_non_zero", True, False), ("obs_exp_lieberman", False, False)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_obs_exp_float_matrices_match_oracle(method, pc, lig, dtype):
    m = random_hic(900, 0.15, 21, integer=False).astype(dtype)
    offs = np.array([0, 310, 640, 900])
    ref = transform_oracle.transform(m, method, offs, per_chromosome=pc, ligation_factor=lig)
    got = obs_exp_transform(m, method, offs, per_chromosome=pc, ligation_factor=lig)
    ref.sort_indices()
    assert got.dtype == ref.dtype
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    # distance sums: fixed-point integer accumulation on the device (order-free, correctly rounded to ~2^-76 of the
    # largest value) against numpy's floating-point sums: rounding-level agreement for float input
    # (one float32 ulp where the reference rounds the quotient to float32)
    rounds_to_f32 = dtype == np.float32 or method == "obs_exp_non_zero"
    np.testing.assert_allclose(got.data, ref.data, rtol=1e-6 if rounds_to_f32 else 1e-12)


def test_hicTransform_cli_obs_exp(tmp_path):
    """mirrors test_hic_transfer_obs_exp / _perChromosome (test_hicTransform.py:20-48)"""
    z = load_golden("normalize_small.npz")
    t = load_golden("transform_small.npz")
    n = int(z["n"])
    up = sparse.csr_matrix((z["one_data"].astype(np.int32), z["one_indices"], z["one_indptr"]), shape=(n, n))
    ma = hiCMatrix()
    ma.matrix = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    ma.cut_intervals = [(c.decode(), int(s), int(e), float(x)) for c, s, e, x in zip(z["chr"], z["start"], z["end"], z["extra"])]
    ma._index_intervals()
    ma.value_dtype = np.int32
    src = str(tmp_path / "small_test_matrix.cool")
    ma.save(src)
    for flags, key in [("", "obs_exp"), ("--perChromosome", "obs_exp_pc")]:
        out = str(tmp_path / ("o" + key + ".cool"))
        main("--matrix {} --outFileName {} --method obs_exp {}".format(src, out, flags).split())
        new = _upper(hiCMatrix(out).matrix)
        np.testing.assert_array_equal(new.data, t["ref_" + key + "_data"])
        np.testing.assert_array_almost_equal(new.data, t["file_" + key + "_data"], decimal=0)   # the reference's own check


def test_distance_sums_are_order_free_and_accurate():
    """integer-limb accumulation: bit-identical between runs, exact for counts, and at least as close to the exact sum
    as numpy's pairwise floating-point sum for float data (also negative values, NaN poisons its distance only)"""
    from fractions import Fraction
    from hicexplorer_b200.store import DeviceCSR
    m = random_hic(1200, 0.2, 5, integer=False)
    m.data[::3] *= -1e-7
    m.data[1::5] *= 1e6
    with DeviceCSR.from_scipy(m) as dev:
        s1, c1 = dev.distance_stats()
        s2, c2 = dev.distance_stats()
    assert np.array_equal(s1, s2) and np.array_equal(c1, c2)
    coo = m.tocoo()
    d = np.abs(coo.row - coo.col)
    assert np.array_equal(c1, np.bincount(d, minlength=1200))
    for dist in (0, 1, 7, 300):
        exact = sum(Fraction(float(v)) for v in coo.data[d == dist])
        assert abs(Fraction(float(s1[dist])) - exact) <= abs(exact) * Fraction(1, 2 ** 50) + Fraction(1, 2 ** 60)
    ints = random_hic(800, 0.2, 6)
    with DeviceCSR.from_scipy(ints) as dev:
        s, _ = dev.distance_stats()
    ci = ints.tocoo()
    np.testing.assert_array_equal(s, np.bincount(np.abs(ci.row - ci.col), weights=ci.data, minlength=800))
    bad = ints.copy()
    bad.data[10] = np.nan
    with DeviceCSR.from_scipy(bad) as dev:
        sb, _ = dev.distance_stats()
    assert np.isnan(sb).sum() == 1


# ---- covariance / pearson (hicTransform.py:105-113, 213-248) ---------------------------------------------------
def _check_against_packed_golden(up, t, key, rtol):
    """the expected matrix is stored as indptr + sha256(indices) + row sums + every 16th cell (make_golden_cov.py)"""
    import hashlib
    assert np.array_equal(up.indptr, t[key + "_indptr"])
    assert hashlib.sha256(up.indices.astype(np.int32).tobytes()).digest() == t[key + "_indices_sha256"].tobytes()
    scale = np.abs(t[key + "_sample"]).max()
    np.testing.assert_allclose(up.data[::int(t["stride"])], t[key + "_sample"], rtol=rtol, atol=rtol * scale)
    np.testing.assert_allclose(np.asarray(up.sum(axis=1)).ravel(), t[key + "_row_sums"], rtol=1e-9, atol=1e-9 * scale)


@pytest.mark.parametrize("method,key", [("covariance", "cov_pc_file"), ("pearson", "pearson_pc_ref")])
def test_correlation_per_chromosome_matches_the_reference_golden(method, key):
    """covariance: the reference's golden FILE test_data/hicTransform/covariance_perChromosome.h5
    (test_hicTransform.py::test_covariance_per_chromosome); pearson: the reference's own _pearson per chromosome"""
    t, _ma, mb = _transform_inputs()
    g = load_golden("transform_cov.npz")
    got = correlation_transform(mb, method, t["b_offsets"], per_chromosome=True)
    assert got.dtype == np.float64 and got.shape == mb.shape
    _check_against_packed_golden(_upper(got), g, key, rtol=1e-11)


@pytest.mark.parametrize("method", ["covariance", "pearson"])
@pytest.mark.parametrize("pc", [False, True])
def test_correlation_matches_oracle(method, pc):
    """random matrices with empty rows (constant variables: pearson's 0/0 -> NaN -> 0) and a one-bin chromosome
    (np.cov's `Degrees of freedom <= 0`: NaN stored by covariance, 0 by pearson)"""
    m = random_hic(700, 0.15, 11, integer=True).tolil()
    for r in (0, 5, 340, 699):
        m[r, :] = 0
        m[:, r] = 0
    m = sparse.csr_matrix(m.tocsr())
    m.eliminate_zeros()
    offs = np.asarray([0, 200, 201, 480, 700], dtype=np.int64)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = transform_oracle.correlation_transform(m, method, offs, per_chromosome=pc)
    got = sparse.csr_matrix(correlation_transform(m, method, offs, per_chromosome=pc))
    got.eliminate_zeros()
    got.sort_indices()
    gd, rd = got.toarray(), ref.toarray()
    assert np.array_equal(np.isnan(gd), np.isnan(rd))
    scale = np.nanmax(np.abs(rd))
    np.testing.assert_allclose(np.nan_to_num(gd), np.nan_to_num(rd), rtol=0, atol=1e-12 * scale)
    if method == "pearson":
        assert np.nanmax(np.abs(gd)) <= 1.0
        live = np.asarray(m.sum(axis=1)).ravel() > 0
        if not pc:
            np.testing.assert_allclose(gd.diagonal()[live], 1.0, atol=1e-12)


def test_hicTransform_cli_covariance(tmp_path):
    """mirrors test_covariance_per_chromosome (test_hicTransform.py: `--method covariance --perChromosome` on
    small_test_matrix_50kb_res.h5 against hicTransform/covariance_perChromosome.h5)"""
    t, _ma, mb = _transform_inputs()
    g = load_golden("transform_cov.npz")
    ma = hiCMatrix()
    ma.matrix = mb
    ma.cut_intervals = [(c.decode(), int(s), int(e), 1.0) for c, s, e in zip(t["b_chr"], t["b_start"], t["b_end"])]
    ma._index_intervals()
    ma.value_dtype = mb.dtype
    src = str(tmp_path / "small_test_matrix_50kb_res.h5")
    ma.save(src)
    out = str(tmp_path / "covariance_perChromosome.h5")
    main("--matrix {} --outFileName {} --method covariance --perChromosome".format(src, out).split())
    _check_against_packed_golden(_upper(hiCMatrix(out).matrix), g, "cov_pc_file", rtol=1e-11)
