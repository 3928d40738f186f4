"""Pin the CPU oracle (oracle/) against the reference's golden vectors and,
where /root/reference is present, against the reference functions themselves."""
import numpy as np
import pytest
from scipy import sparse

from oracle import ice_oracle, kr_oracle, ref_import
from tests.helpers import csr_from_npz, full_from_pixels, random_hic


def test_ice_pipeline_reproduces_reference_golden_file_bit_exact(small_ice):
    z = small_ice
    full = csr_from_npz(z, "in")
    res = ice_oracle.correct_ice_pipeline(full, -1.5, 5.0, max_iter=500)
    assert np.array_equal(res["zero_bins"], z["zero_bins"])
    assert np.array_equal(res["outliers_rel"], z["outliers_rel"])
    assert len(res["outliers_rel"]) == 64  # filtered.bed has 64 lines
    assert res["ice"]["iterations"] == 500  # 59 empty rows pin the deviation at 1
    # the reference's golden output file, bit for bit
    assert np.array_equal(res["upper"].data, z["golden_data"])
    assert np.array_equal(res["upper"].indices, z["golden_indices"])
    assert np.array_equal(res["upper"].indptr, z["golden_indptr"])
    assert np.array_equal(res["nan_bins"], z["golden_nan_bins"])
    np.testing.assert_array_equal(res["correction_factors"], z["golden_correction_factors"])  # NaN-aware


def test_filtered_bed_lists_the_outlier_bins(small_ice):
    z = small_ice
    full = csr_from_npz(z, "in")
    res = ice_oracle.correct_ice_pipeline(full, -1.5, 5.0, max_iter=1)
    want = set()
    for ln in str(z["filtered_bed"]).splitlines():
        c, s, e = ln.split("\t")[:3]
        want.add((c, int(s), int(e)))
    got = set((z["chr"][i].decode(), int(z["start"][i]), int(z["end"][i])) for i in res["outliers_abs"])
    assert got == want


def test_mad_scores_match_reference_values(small_ice, gm12878):
    z = small_ice
    full = csr_from_npz(z, "in")
    keep = np.setdiff1d(np.arange(full.shape[0]), z["zero_bins"])
    st1 = ice_oracle.scrub(ice_oracle.select_bins(full, keep))
    mad = ice_oracle.MadScores(ice_oracle.coverage_points(st1))
    assert mad.median == float(z["mad_median"]) and mad.mad == float(z["mad_mad"])
    np.testing.assert_array_equal(mad.z, z["zscore"])
    g = gm12878
    fullg = full_from_pixels(g)
    keep = np.setdiff1d(np.arange(fullg.shape[0]), g["zero_bins"])
    st1 = ice_oracle.scrub(ice_oracle.select_bins(fullg, keep))
    np.testing.assert_array_equal(ice_oracle.MadScores(ice_oracle.coverage_points(st1)).z, g["zscore"])
    ranges = list(zip(g["chr_lo"].tolist(), g["chr_hi"].tolist()))
    assert np.array_equal(ice_oracle.outlier_bins(st1, -1.5, 5.0, ranges), g["outliers_rel_perchr"])


def test_ice_gm12878_matches_reference_run(gm12878):
    g = gm12878
    res = ice_oracle.correct_ice_pipeline(full_from_pixels(g), -1.5, 5.0, max_iter=500)
    assert np.array_equal(res["zero_bins"], g["zero_bins"])
    assert np.array_equal(res["outliers_rel"], g["outliers_rel"])
    assert res["ice"]["iterations"] == int(g["ice_iterations"]) == 35
    np.testing.assert_array_equal(res["ice"]["bias"], g["ice_bias"])
    w = res["ice"]["matrix"]
    w.sort_indices()
    np.testing.assert_array_equal(w.data[g["ice_sample_idx"]], g["ice_sample_val"])


def test_ice_rejects_asymmetric_input():
    m = random_hic(50, 0.2, 1).tolil()
    m[3, 7] = m[3, 7] + 5
    with pytest.raises(ValueError, match="symmetric"):
        ice_oracle.ice(m.tocsr())


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("seed,n,empty", [(0, 200, 0.0), (1, 333, 0.03), (2, 64, 0.0)])
def test_oracle_equals_reference_functions_on_random_input(seed, n, empty):
    iterativeCorrection, MAD, filter_by_zscore, _ = ref_import.load()
    m = random_hic(n, 0.15, seed, empty_frac=empty, low_frac=0.05)
    W, b = iterativeCorrection(m.copy(), M=40)
    res = ice_oracle.ice(m.copy(), max_iter=40)
    W.sort_indices()
    mine = res["matrix"]
    mine.sort_indices()
    np.testing.assert_array_equal(W.data, mine.data)
    np.testing.assert_array_equal(b, res["bias"])

    class Ma(object):
        matrix = m

    np.testing.assert_array_equal(np.array(filter_by_zscore(Ma, -1.0, 3.0)), ice_oracle.outlier_bins(m, -1.0, 3.0))
    pts = ice_oracle.coverage_points(m)
    np.testing.assert_array_equal(MAD(pts).get_motified_zscores(), ice_oracle.MadScores(pts).z)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_reference_cli_defaults_are_what_we_mirror():
    parse_arguments = ref_import.load()[3]
    a = parse_arguments().parse_args(["correct", "-m", "x.h5", "-o", "y.h5"])
    assert a.correctionMethod == "KR" and a.iterNum == 500 and a.filterThreshold is None
    assert a.perchr is False and a.skipDiagonal is False and a.chromosomes is None


def test_kr_oracle_matches_reference_golden_file(small_ice, small_kr):
    full = csr_from_npz(small_ice, "in")
    res = kr_oracle.kr_pipeline(full, rescale=True)
    up = res["upper"]
    assert np.array_equal(up.indptr, small_kr["golden_indptr"])
    assert np.array_equal(up.indices, small_kr["golden_indices"])
    # the reference's own test pins 5 decimals (test_hicCorrectMatrix.py:52-69)
    np.testing.assert_allclose(up.data, small_kr["golden_data"], rtol=0, atol=1e-7)
    np.testing.assert_array_almost_equal(up.data, small_kr["golden_data"], decimal=5)
    # golden correction_factors follow the older 1/x convention
    np.testing.assert_allclose(1.0 / res["x_out"], small_kr["golden_correction_factors"].ravel(), rtol=2e-6)
    assert res["outer"] == 33 and res["matvecs"] == 252


def test_kr_oracle_matches_kr_partial_weights(kr_partial):
    res = kr_oracle.kr_pipeline(full_from_pixels(kr_partial), rescale=True)
    ratio = kr_partial["weight"] / res["x_out"]
    assert np.ptp(ratio) < 1e-12  # same trajectory
    assert abs(ratio.mean() - 1.0) < 5e-6  # marginally different rescale constant
    assert res["outer"] == 15 and res["matvecs"] == 35


def test_kr_balances_rows():
    m = random_hic(300, 0.2, 5)
    res = kr_oracle.kr_balance(m)
    a = res["A"]
    rows = res["x"] * (a @ res["x"])
    np.testing.assert_allclose(rows, 1.0, atol=2e-6)


def test_kr_small_fixture_is_ill_conditioned(small_ice):
    """Why KR parity on config 1 is checked at 5e-6 and not 1e-9: re-ordering the bins (a pure change of
    floating-point summation order) moves the oracle's own answer by ~1e-6 on this input."""
    full = csr_from_npz(small_ice, "in")
    ref = kr_oracle.kr_pipeline(full, rescale=True)
    perm = np.random.default_rng(0).permutation(full.shape[0])
    alt = kr_oracle.kr_pipeline(full[perm][:, perm].tocsr(), rescale=True)
    x2 = np.empty_like(alt["x_out"])
    x2[perm] = alt["x_out"]
    rel = np.max(np.abs(x2 / ref["x_out"] - 1))
    assert 1e-8 < rel < 5e-6


def test_merge_oracle_reproduces_reference_golden_file():
    from oracle import merge_oracle
    from tests.conftest import load_golden
    z = load_golden("merge_small.npz")
    n = len(z["in_indptr"]) - 1
    up = sparse.csr_matrix((z["in_data"], z["in_indices"], z["in_indptr"]), shape=(n, n))
    full = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    iv = [(c.decode(), int(s), int(e), float(x)) for c, s, e, x in zip(z["in_chr"], z["in_start"], z["in_end"], z["in_extra"])]
    m, new_iv, nan_bins, _map = merge_oracle.merge_bins(full, iv, 5)
    u = sparse.triu(m, 0, format="csr")
    u.sort_indices()
    assert np.array_equal(u.data, z["out_data"]) and np.array_equal(u.indices, z["out_indices"])
    assert np.array_equal(u.indptr, z["out_indptr"]) and np.array_equal(nan_bins, z["out_nan_bins"])
    assert [a[0] for a in new_iv] == [c.decode() for c in z["out_chr"]]
    assert [a[1] for a in new_iv] == z["out_start"].tolist() and [a[2] for a in new_iv] == z["out_end"].tolist()
    np.testing.assert_allclose([a[3] for a in new_iv], z["out_extra"], rtol=1e-12, equal_nan=True)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_merge_oracle_equals_reference_reduce_matrix():
    from oracle import merge_oracle
    ref_import.load()
    from hicexplorer.reduceMatrix import reduce_matrix
    m = random_hic(400, 0.1, 3, empty_frac=0.03)
    iv = [("a" if i < 150 else "b", i * 10, (i + 1) * 10, 1.0) for i in range(400)]
    bin_map, new_iv = merge_oracle.merge_groups(iv, 6)
    groups = [list(np.flatnonzero(bin_map == k)) for k in range(len(new_iv))]
    want = sparse.csr_matrix(reduce_matrix(m, groups, diagonal=True))
    want.eliminate_zeros()
    want.sort_indices()
    got = merge_oracle.reduce(m, bin_map, len(new_iv))
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)


# ---- hicNormalize / hicSumMatrices ---------------------------------------------------------
def _norm_inputs():
    from tests.conftest import load_golden
    z = load_golden("normalize_small.npz")
    n = int(z["n"])

    def full(key):
        up = sparse.csr_matrix((z[key + "_data"], z[key + "_indices"], z[key + "_indptr"]), shape=(n, n))
        m = sparse.csr_matrix(up + sparse.triu(up, 1).T)
        m.sort_indices()
        return m
    return z, n, full


def _upper(m):
    up = sparse.triu(m, 0, format="csr")
    up.sort_indices()
    return up


@pytest.mark.parametrize("mode,keys", [("smallest", ("smallest_one", "smallest_two")),
                                       ("norm_range", ("norm_range_one", "norm_range_two"))])
def test_normalize_oracle_reproduces_reference_golden_files(mode, keys):
    from oracle import normalize_oracle
    z, n, full = _norm_inputs()
    got = normalize_oracle.normalize([full("one"), full("two")], mode)
    for m, key in zip(got, keys):
        up = _upper(m)
        assert up.data.dtype == np.float32
        assert np.array_equal(up.indptr, z[key + "_indptr"]) and np.array_equal(up.indices, z[key + "_indices"])
        np.testing.assert_array_equal(up.data, z[key + "_data"])        # bit-exact float32


def test_normalize_oracle_multiplicative_golden():
    from oracle import normalize_oracle
    z, n, full = _norm_inputs()
    up = _upper(normalize_oracle.normalize([full("one")], "multiplicative", multiplicative_value=2)[0])
    assert np.array_equal(up.indices, z["by2_indices"]) and np.array_equal(up.indptr, z["by2_indptr"])
    np.testing.assert_array_equal(up.data, z["by2_data"])


# ---- hicTransform obs/exp ------------------------------------------------------------------
def _transform_inputs():
    from tests.conftest import load_golden
    z, n, full = _norm_inputs()
    t = load_golden("transform_small.npz")
    nb = int(t["b_n"])
    ub = sparse.csr_matrix((t["b_data"], t["b_indices"], t["b_indptr"]), shape=(nb, nb))
    mb = sparse.csr_matrix(ub + sparse.triu(ub, 1).T)
    mb.sort_indices()
    return t, full("one").astype(np.int32), mb


TRANSFORM_CASES = [("ref_obs_exp", "a", "obs_exp", False, False), ("ref_obs_exp_pc", "a", "obs_exp", True, False),
                   ("ref_non_zero_pc", "a", "obs_exp_non_zero", True, False),
                   ("ref_lieberman_b", "b", "obs_exp_lieberman", False, False),
                   ("ref_non_zero_lig_b", "b", "obs_exp_non_zero", False, True),
                   ("ref_non_zero_lig_pc_b", "b", "obs_exp_non_zero", True, True)]


@pytest.mark.parametrize("key,which,method,pc,lig", TRANSFORM_CASES)
def test_transform_oracle_equals_reference_functions(key, which, method, pc, lig):
    """golden vectors = outputs of the reference's own hicTransform._obs_exp* functions (make_golden_transform.py)"""
    from oracle import transform_oracle
    t, ma, mb = _transform_inputs()
    m, offs = (ma, t["offsets"]) if which == "a" else (mb, t["b_offsets"])
    got = _upper(transform_oracle.transform(m, method, offs, per_chromosome=pc, ligation_factor=lig))
    assert np.array_equal(got.indptr, t[key + "_indptr"]) and np.array_equal(got.indices, t[key + "_indices"])
    np.testing.assert_array_equal(got.data, t[key + "_data"])


@pytest.mark.parametrize("key,file_key", [("ref_obs_exp", "file_obs_exp"), ("ref_obs_exp_pc", "file_obs_exp_pc"),
                                          ("ref_non_zero_pc", "file_non_zero_pc"), ("ref_lieberman_b", "file_lieberman_b")])
def test_transform_reference_golden_files_to_their_own_tolerance(key, file_key):
    """the files under test_data/hicTransform were written by an older code version; the reference's tests compare them
    with assert_array_almost_equal(decimal=0) (test_hicTransform.py:17, 31)"""
    from tests.conftest import load_golden
    t = load_golden("transform_small.npz")
    assert np.array_equal(t[key + "_indices"], t[file_key + "_indices"])
    np.testing.assert_array_almost_equal(t[key + "_data"], t[file_key + "_data"], decimal=0)


def test_correlation_oracle_reproduces_the_reference_golden_file():
    """np.cov per chromosome == test_data/hicTransform/covariance_perChromosome.h5 (same cells; the reference's test
    compares with decimal=0), np.corrcoef per chromosome == the reference's own _pearson (make_golden_cov.py)"""
    import hashlib
    from oracle import transform_oracle
    from tests.conftest import load_golden
    t, _ma, mb = _transform_inputs()
    g = load_golden("transform_cov.npz")
    for method, key in (("covariance", "cov_pc_file"), ("pearson", "pearson_pc_ref")):
        up = _upper(transform_oracle.correlation_transform(mb, method, t["b_offsets"], per_chromosome=True))
        assert np.array_equal(up.indptr, g[key + "_indptr"])
        assert hashlib.sha256(up.indices.astype(np.int32).tobytes()).digest() == g[key + "_indices_sha256"].tobytes()
        np.testing.assert_allclose(up.data[::int(g["stride"])], g[key + "_sample"], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(np.asarray(up.sum(axis=1)).ravel(), g[key + "_row_sums"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("k", [3, 5])
def test_running_window_oracle_equals_the_reference_function(k):
    """oracle/merge_oracle.running_window vs hicMergeMatrixBins.running_window_merge imported from /root/reference"""
    from oracle import merge_oracle, ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    ref_import.load()
    import hicexplorer.hicMergeMatrixBins as ref_mb
    rng = np.random.default_rng(k)
    a = sparse.random(90, 90, 0.12, random_state=rng, format="csr")
    a = sparse.csr_matrix(np.round((a + a.T) * 10)).astype(np.int32)

    class _H(object):
        pass
    h = _H()
    h.matrix = a.copy()
    h.nan_bins = []
    want = sparse.csr_matrix(ref_mb.running_window_merge(h, k).matrix)
    got = merge_oracle.running_window(a, k)
    assert (want != got).nnz == 0 and got.dtype == want.dtype


def test_c_port_on_all_cores_equals_the_serial_c_loop():
    """oracle/hic_cpu.c: the pthread version of the reference's loop (bench.py's all-cores CPU figure) keeps the serial
    summation order, so it equals the single-threaded loop bit for bit"""
    from oracle import hic_cpu
    from tests.helpers import random_hic
    m = random_hic(900, 0.1, 3, integer=False)
    b1, it1, d1, data1 = hic_cpu.ice_loop(m, 12, 0.0, threads=1)
    for threads in (2, 5):
        b, it, d, data, _secs = hic_cpu.ice_loop_mt(m, 12, 0.0, threads=threads)
        assert it == it1 and d == d1
        np.testing.assert_array_equal(b, b1)
        np.testing.assert_array_equal(data, data1)


def test_c_loop_equals_the_numpy_restatement_bit_for_bit():
    """oracle/hic_cpu.c (what the full-size GPU tests and bench.py's CPU legs run) against oracle/ice_oracle.ice, itself
    pinned to the reference: same iteration count, same deviation, same bias to the last bit -- including the mean, which
    numpy forms by pairwise summation (np_pairwise_sum in C restates numpy's @TYPE@_pairwise_sum)."""
    from oracle import hic_cpu
    from tests.helpers import random_hic
    rng = np.random.default_rng(0)
    L = hic_cpu.lib()
    for n in list(range(0, 300)) + [1000, 1023, 1024, 1025, 4097, 12345, 100001]:
        a = rng.random(n) * rng.integers(1, 1000)
        assert L.np_sum_f64(a.ctypes.data, n) == a.sum()
    for seed, (n, dens, integer, empty) in enumerate([(900, 0.1, False, 0.0), (1500, 0.05, True, 0.03), (2500, 0.02, True, 0.0)]):
        m = random_hic(n, dens, seed, integer=integer, empty_frac=empty)
        with np.errstate(divide="ignore"):
            ref = ice_oracle.ice(m.copy(), max_iter=60, tolerance=1e-5)
        for run in (lambda: hic_cpu.ice_loop(m, 60, 1e-5, threads=1)[:3], lambda: hic_cpu.ice_loop_mt(m, 60, 1e-5, threads=3)[:3]):
            b, it, d = run()
            assert it == ref["iterations"] and d == ref["deviation"]
            np.testing.assert_array_equal(b / b[b != 0].mean(), ref["bias"])
        # the masked CSR form: iterativeCorrection on matrix[keep][:, keep] without building the sub-matrix
        keep = rng.random(n) > 0.1
        sub = m[keep][:, keep].tocsr()
        sub.sort_indices()
        with np.errstate(divide="ignore"):
            ref2 = ice_oracle.ice(sub.copy(), max_iter=60, tolerance=1e-5)
        for threads in (1, 4):
            b2, it2, d2, data2, _s = hic_cpu.ice_loop_masked(m, keep.astype(np.uint8), 60, 1e-5, threads=threads)
            assert it2 == ref2["iterations"] and d2 == ref2["deviation"]
            np.testing.assert_array_equal(b2 / b2[b2 != 0].mean(), ref2["bias"])
        rs, dg = hic_cpu.row_stats_masked(m, keep.astype(np.uint8), threads=3)
        # row sums: exact for counts (any order); scipy's own summation order is not restated for non-integer values
        (np.testing.assert_array_equal if integer else np.testing.assert_allclose)(rs[keep], np.asarray(sub.sum(axis=1)).ravel())
        np.testing.assert_array_equal(dg[keep], sub.diagonal())


def test_threaded_kr_operator_equals_the_scipy_restatement():
    from oracle import kr_oracle
    from tests.helpers import random_hic
    for seed in range(2):
        m = random_hic(1200, 0.05, seed, low_frac=0.05)
        a = kr_oracle.kr_balance(m)
        b = kr_oracle.kr_balance(m, threaded=True)
        assert (a["outer"], a["matvecs"]) == (b["outer"], b["matvecs"])
        np.testing.assert_allclose(b["x"], a["x"], rtol=1e-13)
