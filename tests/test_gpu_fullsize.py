"""Size-independent properties checked at BASELINE.json's sizes, where the CPU oracle would take minutes to hours:
the synthetic 25 kb matrix of configs[1] (~3e8 stored entries) and the 10 kb matrix of configs[2] (~1.5e9).

  * the generator is exactly symmetric,
  * ICE converges and the corrected matrix is balanced: every non-empty row sums to the common mean (Imakaev's
    defining property; what `deviation < tolerance` certifies, iterativeCorrection.py:47,72),
  * idempotence: correcting the corrected matrix converges immediately with a flat bias,
  * KR: x o ((A + 1e-5 I) x) = 1 to the KR tolerance, x > 0,
  * the MAD filter finds the planted low-coverage bins and the zero-bin mask the planted empty bins."""
import ctypes

import numpy as np
import pytest

import bench
from hicexplorer_b200 import _lib
from hicexplorer_b200.store import DeviceCSR

pytestmark = pytest.mark.gpu


def _row_sums_of_scaled(dev, scale):
    """r_i * sum_j A_ij r_j for r = scale, through the same SpMV kernel (hb_dev_spmv)."""
    import torch
    n = dev.n
    r = torch.from_numpy(np.ascontiguousarray(scale)).cuda()
    out = torch.zeros(n, dtype=torch.float64, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(_lib.lib().hb_dev_spmv(dev.handle, P(r), 0.0, P(r), None, None, P(out), 0, None))
    dev.sync()
    return out.cpu().numpy()


@pytest.mark.parametrize("workload", ["ice25k", "ice10k"])
def test_ice_balances_the_matrix_at_full_size(workload):
    method, bins, kw = bench.synth_kwargs(workload, clean=True)
    with DeviceCSR.synthetic(bins, **kw) as dev:
        assert dev.nnz > 0.95 * bench.WORKLOADS[workload][2]
        assert dev.symmetry_ratio() == 0.0                       # generator: exactly symmetric
        bias, iters, devn = dev.ice(500, 1e-5)
        assert devn[0] < 1e-5 and 10 < int(iters[0]) < 500       # converged
        assert np.all(bias > 0) and abs(bias.mean() - 1.0) < 1e-12
        rows = _row_sums_of_scaled(dev, 1.0 / bias)              # row sums of A_ij / (b_i b_j)
        assert np.max(np.abs(rows / rows.mean() - 1.0)) < 3e-5   # balanced
        # idempotence: one more ICE pass on the balanced matrix changes nothing beyond the tolerance
        values = dev.ice_scaled_values()
        assert np.isfinite(values).all() and values.min() > 0
    del values


def test_ice_is_idempotent_on_its_own_output():
    method, bins, kw = bench.synth_kwargs("ice100k", clean=True)
    with DeviceCSR.synthetic(bins, **kw) as dev:
        bias, iters, devn = dev.ice(500, 1e-5)
        values = dev.ice_scaled_values()
        pattern = dev.download()
    assert devn[0] < 1e-5
    pattern.data[:] = values
    with DeviceCSR.from_scipy(pattern) as dev2:
        bias2, iters2, devn2 = dev2.ice(500, 1e-5)
    assert int(iters2[0]) <= 2 and np.max(np.abs(bias2 - 1.0)) < 5e-5


def test_kr_balances_the_matrix_at_full_size():
    method, bins, kw = bench.synth_kwargs("kr25k", clean=False)
    with DeviceCSR.synthetic(bins, **kw) as dev:
        x, outer, matvecs, residual = dev.kr()
        assert residual < 1e-6 and x.min() > 0 and 5 < outer < 200
        import torch
        n = dev.n
        xt = torch.from_numpy(x).cuda()
        out = torch.zeros(n, dtype=torch.float64, device="cuda")
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        _lib.check(_lib.lib().hb_dev_spmv(dev.handle, P(xt), 1e-5, P(xt), None, None, P(out), 0, None))
        dev.sync()
        v = out.cpu().numpy()
        assert np.max(np.abs(v - 1.0)) < 1e-5                    # x o ((A + eps I) x) = 1
        rs, dg = dev.row_stats()
        empty = rs == 0
        assert 0.002 < empty.mean() < 0.01                       # the planted 0.5 % empty bins ...
        np.testing.assert_allclose(x[empty], 1.0 / np.sqrt(1e-5), rtol=1e-6)   # ... balance against the 1e-5 diagonal


def test_filter_finds_planted_bins_at_full_size():
    method, bins, kw = bench.synth_kwargs("ice25k", clean=False)
    with DeviceCSR.synthetic(bins, **kw) as dev:
        n = dev.n
        rs, dg = dev.row_stats()
        zero = rs == 0
        assert 0.002 < zero.mean() < 0.01
        assert np.all(rs == np.floor(rs))                        # integer counts: row sums exact in any order
        dev.compact(zero.astype(np.uint8))
        mask, z, med, mad = dev.mad_filter(-0.85, 50.0, want_z=True)
        pts = (rs - dg)[~zero]
        # same statistics with numpy on the downloaded O(n) vector (the O(nnz) part stayed on the GPU)
        med_ref = np.median(pts[pts > 0])
        mad_ref = np.median(np.abs(pts - med_ref))
        assert med[0] == med_ref and mad[0] == mad_ref
        zr = 0.6745 * ((pts - med_ref) / mad_ref)
        np.testing.assert_array_equal(z, zr)                     # bit-exact z-scores at 123 k bins
        np.testing.assert_array_equal(mask.astype(bool), (zr < -0.85) | (zr > 50.0))
        from oracle.synth_ref import bin_bias
        g = bin_bias(kw["seed"], n, kw["low_frac"], kw["empty_frac"])
        assert np.array_equal(g == 0.0, zero)                    # the zero-bin mask is exactly the planted empty bins
        planted_low = (g == 0.02)[~zero]
        assert np.all(mask[planted_low] == 1)                    # every planted low-coverage bin is an outlier
        assert 0.015 < mask.mean() < 0.06
        dev.compact(mask)
        assert dev.n == n - zero.sum() - mask.sum()
        bias, iters, devn = dev.ice(500, 1e-5)
        assert devn[0] < 1e-5


def test_genome_wide_100kb_pipeline_matches_the_reference_run():
    """4.0e7-entry genome-wide matrix (far/trans entries, planted empty and low-coverage bins) against what the
    REFERENCE's own functions produced for it (tests/golden/make_golden_synth100k.py): masks bit-exact, same iteration
    count +-1, bias 1e-9, corrected values 1e-8"""
    from hicexplorer_b200.pipeline import ice_pipeline
    from tests.conftest import load_golden
    z = load_golden("synth100k_ice.npz")
    method, bins, kw = bench.synth_kwargs("ice100k", clean=False)
    with DeviceCSR.synthetic(bins, **kw) as dev:
        m = dev.download()
    assert m.shape[0] == int(z["n"])
    got = ice_pipeline(m, (float(z["lo"]), float(z["hi"])), max_iter=500)
    assert np.array_equal(got["zero_bins"], z["zero_bins"])
    assert np.array_equal(got["outliers_rel"], z["outliers_rel"])            # bit-exact filter mask
    assert np.array_equal(got["kept"], z["kept"])
    assert abs(int(got["iterations"][0]) - int(z["iterations"])) <= 1
    np.testing.assert_allclose(got["bias"], z["bias"], rtol=1e-9)
    w = got["matrix"]
    assert w.nnz == int(z["nnz"])
    np.testing.assert_allclose(w.data[z["sample_idx"]], z["sample_val"], rtol=1e-8)
    np.testing.assert_allclose(np.asarray(w.sum(axis=1)).ravel(), z["row_sums"], rtol=1e-8)
    np.testing.assert_allclose(w.data.sum(), float(z["data_sum"]), rtol=1e-9)


# ---------------------------------------------------------------------------------------------------------------
# Oracle parity AT BASELINE.json's sizes.  The checker is oracle/hic_cpu.c on every host core: the reference's loop
# (row sums, mean over non-empty rows with numpy's pairwise summation, two in-place rescaling passes, 1e100 scan,
# strict '<' test) on matrix[keep][:, keep] without building the sub-matrix -- pinned bit for bit to
# oracle/ice_oracle.ice, itself pinned to the reference (tests/test_oracle.py).  Input: the synthetic matrix WITH the
# planted empty and low-coverage bins, generated on the device and downloaded, so that both masks do real work.
def _host_memory_ok(nnz):
    try:
        import psutil
        return psutil.virtual_memory().available > 2.2 * 12.0 * nnz + 8e9
    except Exception:
        return True


def _oracle_ice_pipeline_fullsize(m, lo, hi, max_iter=500, tol=1e-5):
    from oracle import hic_cpu, ice_oracle
    n = m.shape[0]
    rs, _dg = hic_cpu.row_stats_masked(m)                       # hicCorrectMatrix.py:612-617 (diagonal included)
    zero = rs == 0
    keep1 = ~zero
    rs1, dg1 = hic_cpu.row_stats_masked(m, keep1.astype(np.uint8))
    pts = (rs1 - dg1)[keep1]                                    # :584-588
    out = ice_oracle.MadScores(pts).outliers(lo, hi)            # :349-391
    keep2 = keep1.copy()
    keep2[np.flatnonzero(keep1)[out]] = False
    bias, iters, dev, _data, secs = hic_cpu.ice_loop_masked(m, keep2.astype(np.uint8), max_iter, tol, inplace=True)
    bias = bias / bias[bias != 0].mean()                        # iterativeCorrection.py:77-78
    return dict(zero_bins=np.flatnonzero(zero), outliers_rel=np.flatnonzero(out), kept=np.flatnonzero(keep2), bias=bias,
                iterations=iters, deviation=dev, seconds=secs, n=n)


@pytest.mark.parametrize("workload", ["ice25k", "ice10k"])
@pytest.mark.timeout(1500)
def test_ice_pipeline_equals_the_oracle_at_full_size(workload):
    """configs[1]'s matrix (3.0e8 entries) and configs[2]'s (1.5e9): both masks exactly, iterations +-1, bias 1e-9"""
    from hicexplorer_b200.pipeline import ice_pipeline
    method, bins, kw = bench.synth_kwargs(workload, clean=False)
    if not _host_memory_ok(bench.WORKLOADS[workload][2]):
        pytest.skip("not enough host memory for the CPU oracle at this size")
    with DeviceCSR.synthetic(bins, **kw) as dev:
        m = dev.download()
    got = ice_pipeline(m, (-0.85, 50.0), max_iter=500, want_matrix=False)
    ref = _oracle_ice_pipeline_fullsize(m, -0.85, 50.0)          # overwrites m.data with the oracle's scaled values
    assert len(ref["zero_bins"]) > 0 and len(ref["outliers_rel"]) > 0
    np.testing.assert_array_equal(got["zero_bins"], ref["zero_bins"])         # bit-exact masks
    np.testing.assert_array_equal(got["outliers_rel"], ref["outliers_rel"])
    np.testing.assert_array_equal(got["kept"], ref["kept"])
    assert ref["deviation"] < 1e-5
    assert abs(int(got["iterations"][0]) - ref["iterations"]) <= 1
    np.testing.assert_allclose(got["bias"], ref["bias"], rtol=1e-9)
    print("%s: %d bins, %d entries, %d / %d iterations (GPU / oracle), oracle loop %.1f s, max rel bias error %.2e"
          % (workload, m.shape[0], m.nnz, int(got["iterations"][0]), ref["iterations"], ref["seconds"],
             float(np.max(np.abs(got["bias"] / ref["bias"] - 1.0)))))


@pytest.mark.timeout(900)
def test_kr_equals_the_oracle_at_full_size():
    """configs[1]: KR on the 25 kb matrix (3.0e8 entries, planted empty / low-coverage bins) against the restated bnewt
    driven by the row-parallel C mat-vec: same outer / mat-vec counts, x within 1e-9"""
    from hicexplorer_b200.pipeline import kr_pipeline
    from oracle import kr_oracle
    method, bins, kw = bench.synth_kwargs("kr25k", clean=False)
    with DeviceCSR.synthetic(bins, **kw) as dev:
        m = dev.download()
    got = kr_pipeline(m, rescale=False, want_matrix=False)
    ref = kr_oracle.kr_balance(m, threaded=True)
    assert (got["outer"], got["matvecs"]) == (ref["outer"], ref["matvecs"])
    np.testing.assert_allclose(got["x"], ref["x"], rtol=1e-9)
    print("kr25k: %d bins, %d entries, %d outer / %d mat-vecs, max rel x error %.2e"
          % (m.shape[0], m.nnz, got["outer"], got["matvecs"], float(np.max(np.abs(got["x"] / ref["x"] - 1.0)))))
