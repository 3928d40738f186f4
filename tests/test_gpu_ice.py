"""GPU parity tests for ICE and the bin filter: CUDA path (through the C-ABI) vs the CPU oracle and the
committed golden vectors.  Tolerances are the north star's: mask bit-exact; bias 1e-9 relative; corrected
matrix 1e-8 relative; iteration count +-1."""
import numpy as np
import pytest
from scipy import sparse

from oracle import ice_oracle
from tests.helpers import csr_from_npz, full_from_pixels, random_hic

pytestmark = pytest.mark.gpu

BIAS_RTOL = 1e-9
DATA_RTOL = 1e-8


def _close(a, b, rtol):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=0)


def test_config1_pipeline_matches_reference_golden(small_ice):
    from hicexplorer_b200.pipeline import ice_pipeline, scatter_to_frame
    z = small_ice
    full = csr_from_npz(z, "in")
    res = ice_pipeline(full, (-1.5, 5.0), max_iter=500)
    assert np.array_equal(res["zero_bins"], z["zero_bins"])          # bit-exact masks
    assert np.array_equal(res["outliers_rel"], z["outliers_rel"])
    assert int(res["iterations"][0]) == 500                           # empty rows pin the deviation at 1
    _close(res["bias"], z["ice_bias"], BIAS_RTOL)
    assert np.array_equal(res["bias"] == 0, z["ice_bias"] == 0)       # 59 dead bins -> exactly 0
    w = res["matrix"]
    assert np.array_equal(w.indices, z["ice_out_indices"]) and np.array_equal(w.indptr, z["ice_out_indptr"])
    _close(w.data, z["ice_out_data"], DATA_RTOL)
    # the file the reference writes: upper triangle in the original frame, NaN-padded factors
    up = sparse.triu(scatter_to_frame(w, res["kept"], res["n"]), 0, format="csr")
    up.eliminate_zeros()
    up.sort_indices()
    assert np.array_equal(up.indices, z["golden_indices"]) and np.array_equal(up.indptr, z["golden_indptr"])
    _close(up.data, z["golden_data"], DATA_RTOL)
    cf = np.full(res["n"], np.nan)
    cf[res["kept"]] = res["bias"]
    assert np.array_equal(np.isnan(cf), np.isnan(z["golden_correction_factors"]))
    _close(cf[res["kept"]], z["golden_correction_factors"][res["kept"]], BIAS_RTOL)
    assert np.array_equal(np.flatnonzero(np.isnan(cf)), z["golden_nan_bins"])


def test_gm12878_pipeline_matches_reference_run(gm12878):
    from hicexplorer_b200.pipeline import ice_pipeline
    g = gm12878
    res = ice_pipeline(full_from_pixels(g), (-1.5, 5.0), max_iter=500)
    assert np.array_equal(res["zero_bins"], g["zero_bins"])
    assert np.array_equal(res["outliers_rel"], g["outliers_rel"])
    assert abs(int(res["iterations"][0]) - int(g["ice_iterations"])) <= 1
    _close(res["bias"], g["ice_bias"], BIAS_RTOL)
    _close(res["matrix"].data[g["ice_sample_idx"]], g["ice_sample_val"], DATA_RTOL)
    _close(np.asarray(res["matrix"].sum(axis=1)).ravel(), g["ice_rowsum"], 1e-9)


def test_gm12878_perchr_filter_mask_bit_exact(gm12878):
    from hicexplorer_b200.store import DeviceCSR
    g = gm12878
    full = full_from_pixels(g)
    keep = np.setdiff1d(np.arange(full.shape[0]), g["zero_bins"])
    st1 = ice_oracle.select_bins(full, keep)
    with DeviceCSR.from_scipy(st1) as dev:
        dev.set_segments(np.concatenate([g["chr_lo"], g["chr_hi"][-1:]]))
        mask, z, med, mad = dev.mad_filter(-1.5, 5.0, want_z=True)
    assert np.array_equal(np.flatnonzero(mask), g["outliers_rel_perchr"])


def test_zscores_bit_exact(small_ice, gm12878):
    from hicexplorer_b200.store import DeviceCSR
    for z, full in ((small_ice, csr_from_npz(small_ice, "in")), (gm12878, full_from_pixels(gm12878))):
        keep = np.setdiff1d(np.arange(full.shape[0]), z["zero_bins"])
        st1 = ice_oracle.select_bins(full, keep)
        with DeviceCSR.from_scipy(st1) as dev:
            mask, zs, med, mad = dev.mad_filter(-1.5, 5.0, want_z=True)
        np.testing.assert_array_equal(zs, z["zscore"])
        ref = ice_oracle.MadScores(ice_oracle.coverage_points(st1))
        assert med[0] == ref.median and mad[0] == ref.mad


@pytest.mark.parametrize("seed,n,density,empty", [(0, 257, 0.2, 0.0), (1, 1000, 0.05, 0.02), (2, 33, 0.5, 0.0),
                                                  (3, 4099, 0.01, 0.01), (4, 1, 1.0, 0.0)])
def test_drop_in_function_matches_oracle_on_random_input(seed, n, density, empty):
    from hicexplorer_b200.iterativeCorrection import iterativeCorrection
    m = random_hic(n, density, seed, empty_frac=empty, low_frac=0.03)
    ref = ice_oracle.ice(m.copy(), max_iter=60, tolerance=1e-5)
    w, bias = iterativeCorrection(m.copy(), M=60, tolerance=1e-5)
    assert isinstance(w, sparse.csr_matrix) and w.dtype == np.float64 and bias.shape == (n,)
    r = ref["matrix"]
    r.sort_indices()
    assert np.array_equal(w.indices, r.indices) and np.array_equal(w.indptr, r.indptr)
    _close(bias, ref["bias"], BIAS_RTOL)
    _close(w.data, r.data, DATA_RTOL)


def test_iteration_count_and_tolerance_semantics():
    from hicexplorer_b200.store import DeviceCSR
    m = random_hic(600, 0.1, 7)
    for tol in (1e-2, 1e-5, 1e-9):
        ref = ice_oracle.ice(m.copy(), max_iter=400, tolerance=tol)
        with DeviceCSR.from_scipy(m) as dev:
            bias, iters, dev_last = dev.ice(400, tol)
        assert abs(int(iters[0]) - ref["iterations"]) <= 1
        assert (dev_last[0] < tol) == (ref["deviation"] < tol)
        np.testing.assert_allclose(dev_last[0], ref["deviation"], rtol=1e-4, atol=1e-15)
        _close(bias, ref["bias"], BIAS_RTOL if tol <= 1e-5 else 1e-6)
    with DeviceCSR.from_scipy(m) as dev:      # M caps the loop; M = 0 leaves the bias at 1
        bias, iters, _ = dev.ice(3, 0.0)
        assert int(iters[0]) == 3
        bias0, iters0, _ = dev.ice(0, 1e-5)
        assert int(iters0[0]) == 0 and np.all(bias0 == 1.0)


def test_non_canonical_and_non_csr_inputs():
    from hicexplorer_b200.iterativeCorrection import iterativeCorrection
    m = random_hic(300, 0.1, 11)
    ref = ice_oracle.ice(m.copy(), max_iter=30)
    for variant in (m.tocoo(), m.tocsc(), m.astype(np.int64), m.tolil()):
        w, bias = iterativeCorrection(variant, M=30)
        _close(bias, ref["bias"], BIAS_RTOL)
    # unsorted indices
    u = m.copy()
    for i in range(u.shape[0]):
        s, e = u.indptr[i], u.indptr[i + 1]
        u.indices[s:e] = u.indices[s:e][::-1]
        u.data[s:e] = u.data[s:e][::-1]
    u.has_sorted_indices = False
    w, bias = iterativeCorrection(u, M=30)
    _close(bias, ref["bias"], BIAS_RTOL)


def test_nan_input_is_scrubbed_in_the_callers_matrix():
    from hicexplorer_b200.iterativeCorrection import iterativeCorrection
    m = random_hic(120, 0.2, 3)
    i, j = m.nonzero()
    k = np.flatnonzero(i < j)[5]
    m[i[k], j[k]] = np.nan
    m[j[k], i[k]] = np.nan
    ref_in = m.copy()
    ref = ice_oracle.ice(ref_in, max_iter=25)
    w, bias = iterativeCorrection(m, M=25)
    assert not np.isnan(m.data).any()          # the reference mutates the caller's data (:28-30)
    _close(bias, ref["bias"], BIAS_RTOL)


def test_asymmetric_matrix_raises_value_error():
    from hicexplorer_b200.iterativeCorrection import iterativeCorrection
    m = random_hic(80, 0.3, 5).tolil()
    m[3, 9] = m[3, 9] + 4
    with pytest.raises(ValueError, match="Please provide symmetric matrix!"):
        iterativeCorrection(m.tocsr())
    # an entry present on one side only
    m = random_hic(80, 0.3, 6).tolil()
    m[2, 70] = 3.0
    m[70, 2] = 0.0
    with pytest.raises(ValueError, match="symmetric"):
        iterativeCorrection(m.tocsr())


def test_overflow_guards_exit_like_the_reference():
    from hicexplorer_b200.iterativeCorrection import iterativeCorrection
    # a bin that only touches itself weakly, next to heavy bins, blows up the final 1e10 guard
    n = 6
    d = np.full((n, n), 1e6)
    d[0, :] = 0
    d[:, 0] = 0
    d[0, 0] = 1e-9
    m = sparse.csr_matrix(d)
    with pytest.raises(ice_oracle.IceOverflow):
        ice_oracle.ice(m.copy(), max_iter=5)
    with pytest.raises(SystemExit) as ei:
        iterativeCorrection(m.copy(), M=5)
    assert ei.value.code == 1


def test_perchr_segments_equal_independent_runs():
    from hicexplorer_b200.pipeline import ice_pipeline
    blocks = [random_hic(n, 0.15, 20 + k) for k, n in enumerate((150, 90, 1200, 64))]
    full = sparse.block_diag(blocks, format="csr").tolil()
    full[10, 200] = full[200, 10] = 7.0      # trans contacts: must be ignored and dropped (perchr)
    full[5, 1300] = full[1300, 5] = 3.0
    full = full.tocsr()
    offs = np.concatenate([[0], np.cumsum([b.shape[0] for b in blocks])])
    res = ice_pipeline(full, (-1e6, 1e6), max_iter=200, chr_offsets=offs, perchr=True)
    assert len(res["outliers_rel"]) == 0 and len(res["zero_bins"]) == 0
    for k, b in enumerate(blocks):
        ref = ice_oracle.ice(b.copy(), max_iter=200)
        a, e = offs[k], offs[k + 1]
        _close(res["bias"][a:e], ref["bias"], BIAS_RTOL)
        assert abs(int(res["iterations"][k]) - ref["iterations"]) <= 1
        sub = res["matrix"][a:e, a:e]
        r = ref["matrix"]
        r.sort_indices()
        _close(sub.data, r.data, DATA_RTOL)
    assert res["matrix"].nnz == sum(b.nnz for b in blocks)      # trans entries dropped


def test_verbose_mode_reports_like_the_reference(caplog):
    import logging
    from hicexplorer_b200.iterativeCorrection import iterativeCorrection
    m = random_hic(200, 0.2, 9)
    ref = ice_oracle.ice(m.copy(), max_iter=100, tolerance=0.05)
    assert ref["deviation"] < 0.05 and 5 < ref["iterations"] < 100
    with caplog.at_level(logging.INFO, logger="hicexplorer_b200.iterativeCorrection"):
        w, bias = iterativeCorrection(m, M=100, tolerance=0.05, verbose=True)
    text = caplog.text
    assert "starting iterative correction" in text and "pass 5 Estimated time" in text and "max delta - 1 = " in text
    # the reference logs loop bodies + 1 (iterativeCorrection.py:41,73)
    assert "[iterative correction] {} iterations used".format(ref["iterations"] + 1) in text
    _close(bias, ref["bias"], 1e-6)
    # without convergence nothing is logged, like the reference
    caplog.clear()
    with caplog.at_level(logging.INFO, logger="hicexplorer_b200.iterativeCorrection"):
        iterativeCorrection(m, M=3, tolerance=1e-12, verbose=True)
    assert "iterations used" not in caplog.text


def test_thousands_of_segments():
    from hicexplorer_b200.store import DeviceCSR
    """draft assemblies have thousands of contigs: --perchr with more than 4096 independent blocks"""
    from oracle import ice_oracle
    rng = np.random.default_rng(12)
    sizes = rng.integers(3, 8, size=5000)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    blocks = []
    for s in sizes:
        b = rng.integers(1, 50, size=(s, s)).astype(np.float64)
        blocks.append(sparse.csr_matrix(b + b.T))
    m = sparse.block_diag(blocks, format="csr")
    m.sort_indices()
    with DeviceCSR.from_scipy(m) as dev:
        dev.set_segments(offs)
        bias, iters, devn = dev.ice(200, 1e-5)
    assert len(iters) == 5000
    for k in rng.choice(5000, size=40, replace=False):
        ref = ice_oracle.ice(blocks[k].copy(), max_iter=200, tolerance=1e-5)
        np.testing.assert_allclose(bias[offs[k]:offs[k + 1]], ref["bias"], rtol=1e-9)
        assert abs(int(iters[k]) - ref["iterations"]) <= 1
