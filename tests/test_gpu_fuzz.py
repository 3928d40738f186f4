"""Randomised cross-checks of the ICE pipeline: every route through it (full or upper-triangle upload, kept-frame or
file-layout result, chromosome subset deleted on the device, --perchr, --skipDiagonal) against the CPU oracle and against
each other, on seeded random matrices of varying size, density and defect rates."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200.pipeline import ice_pipeline, scatter_to_frame
from oracle import ice_oracle
from tests.helpers import random_hic

pytestmark = pytest.mark.gpu


def _oracle(full, lo, hi, offs, perchr, skip_diagonal, max_iter):
    """the driver's ICE branch restated with the oracle's pieces (hicCorrectMatrix.py:612-740)"""
    n = full.shape[0]
    dead = ice_oracle.zero_bins(full)
    keep1 = np.setdiff1d(np.arange(n), dead)
    stage = ice_oracle.scrub(ice_oracle.select_bins(full, keep1))
    if skip_diagonal:
        stage = stage.tolil()
        stage.setdiag(0)
        stage = stage.tocsr()
        stage.eliminate_zeros()
    offs1 = np.searchsorted(keep1, offs, side="left")
    ranges = list(zip(offs1[:-1], offs1[1:])) if perchr else None
    out_rel = ice_oracle.outlier_bins(stage.copy(), lo, hi, chr_ranges=ranges)
    keep2 = np.delete(np.arange(stage.shape[0]), out_rel)
    ice_in = ice_oracle.select_bins(stage, keep2)
    kept = keep1[keep2]
    if not perchr:
        res = ice_oracle.ice(ice_in.copy(), max_iter=max_iter)
        return dead, out_rel, kept, res["bias"], sparse.csr_matrix(res["matrix"])
    offs2 = np.searchsorted(kept, offs, side="left")
    blocks, bias = [], []
    for a, b in zip(offs2[:-1], offs2[1:]):
        r = ice_oracle.ice(sparse.csr_matrix(ice_in[a:b, a:b]), max_iter=max_iter)
        blocks.append(sparse.csr_matrix(r["matrix"]))
        bias.append(r["bias"])
    return dead, out_rel, kept, np.concatenate(bias), sparse.block_diag(blocks, format="csr")


@pytest.mark.parametrize("seed", range(16))
def test_pipeline_routes_agree_with_the_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(60, 1400))
    density = float(rng.choice([0.02, 0.08, 0.3, 0.9]))
    full = random_hic(n, density, seed, empty_frac=float(rng.choice([0.0, 0.03])), low_frac=float(rng.choice([0.0, 0.05])),
                      integer=bool(rng.integers(0, 2)))
    nchr = int(rng.integers(1, 6))
    offs = np.concatenate([[0], np.sort(rng.choice(np.arange(1, n), nchr - 1, replace=False)), [n]]).astype(np.int64)
    perchr = bool(rng.integers(0, 2)) and nchr > 1
    skip_diagonal = bool(rng.integers(0, 2))
    lo, hi = float(rng.choice([-1.5, -3.0, -1e6])), float(rng.choice([5.0, 2.5, 1e6]))
    max_iter = int(rng.choice([30, 200]))
    kw = dict(max_iter=max_iter, chr_offsets=offs, perchr=perchr, skip_diagonal=skip_diagonal)
    try:
        dead, out_rel, kept, bias, w = _oracle(full, lo, hi, offs, perchr, skip_diagonal, max_iter)
    except ice_oracle.IceOverflow:
        # the reference exits here (iterativeCorrection.py:58-62, 80-84): the device path must stop the same way
        from hicexplorer_b200 import _lib
        with pytest.raises((_lib.OverflowIter, _lib.OverflowFinal)):
            ice_pipeline(full, (lo, hi), **kw)
        return
    a = ice_pipeline(full, (lo, hi), **kw)
    up = sparse.triu(full, 0, format="csr")
    b = ice_pipeline(up, (lo, hi), upper_input=True, file_layout=True, **kw)
    for got in (a, b):
        assert np.array_equal(got["zero_bins"], dead)
        assert np.array_equal(got["outliers_rel"], out_rel)              # masks bit-exact
        assert np.array_equal(got["kept"], kept)
        ok = np.isfinite(bias)
        np.testing.assert_allclose(got["bias"][ok], bias[ok], rtol=1e-9, atol=1e-300)
    np.testing.assert_array_equal(a["bias"], b["bias"])                  # the two upload routes: same bits
    w.sort_indices()
    assert a["matrix"].nnz == w.nnz
    good = np.isfinite(w.data)
    np.testing.assert_allclose(a["matrix"].data[good], w.data[good], rtol=1e-8)
    want_up = sparse.triu(scatter_to_frame(a["matrix"], a["kept"], n), 0, format="csr")
    want_up.eliminate_zeros()
    want_up.sort_indices()
    assert np.array_equal(b["upper"].indptr, want_up.indptr) and np.array_equal(b["upper"].indices, want_up.indices)
    np.testing.assert_array_equal(b["upper"].data, want_up.data)         # file layout == kept-frame result, same bits
    # a chromosome subset deleted on the device == the same pipeline on the pre-sliced matrix
    if nchr > 2:
        chosen = np.sort(rng.choice(nchr, size=nchr - 1, replace=False))
        keep_bins = np.concatenate([np.arange(offs[c], offs[c + 1]) for c in chosen])
        drop = np.ones(n, dtype=np.uint8)
        drop[keep_bins] = 0
        sub = sparse.csr_matrix(full[keep_bins, :][:, keep_bins])
        sub.sort_indices()
        sub_offs = np.concatenate([[0], np.cumsum([offs[c + 1] - offs[c] for c in chosen])]).astype(np.int64)
        kw2 = dict(kw, chr_offsets=sub_offs)
        c = ice_pipeline(up, (lo, hi), upper_input=True, file_layout=True, drop_mask=drop, **kw2)
        d = ice_pipeline(sub, (lo, hi), file_layout=True, **kw2)
        assert np.array_equal(c["kept"], d["kept"])
        np.testing.assert_array_equal(c["bias"], d["bias"])
        assert np.array_equal(c["upper"].indices, d["upper"].indices)
        np.testing.assert_array_equal(c["upper"].data, d["upper"].data)


@pytest.mark.parametrize("seed", range(6))
def test_kr_routes_agree(seed):
    """Knight-Ruiz through the full upload, the upper-triangle upload and a device-side chromosome subset: identical bits;
    against the oracle 1e-9 (x) / 1e-8 (matrix) on matrices without empty bins"""
    from hicexplorer_b200.pipeline import kr_pipeline
    from oracle import kr_oracle
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.integers(80, 900))
    full = random_hic(n, float(rng.choice([0.1, 0.4])), 50 + seed, integer=bool(rng.integers(0, 2)))
    up = sparse.triu(full, 0, format="csr")
    a = kr_pipeline(full, rescale=True)
    b = kr_pipeline(up, rescale=True, upper_input=True)
    np.testing.assert_array_equal(a["x"], b["x"])
    np.testing.assert_array_equal(a["upper"].data, b["upper"].data)
    assert a["matvecs"] == b["matvecs"] and a["factor"] == b["factor"]
    ref = kr_oracle.kr_pipeline(full, rescale=True)
    np.testing.assert_allclose(a["x"], ref["x_out"], rtol=1e-9)
    np.testing.assert_allclose(a["upper"].data, ref["upper"].data, rtol=1e-8)
    cut = int(rng.integers(n // 4, 3 * n // 4))
    drop = np.zeros(n, dtype=np.uint8)
    drop[cut:] = 1
    c = kr_pipeline(up, rescale=True, upper_input=True, drop_mask=drop)
    d = kr_pipeline(sparse.csr_matrix(full[:cut, :cut]), rescale=True)
    np.testing.assert_array_equal(c["x"], d["x"])
    np.testing.assert_array_equal(c["upper"].data, d["upper"].data)
