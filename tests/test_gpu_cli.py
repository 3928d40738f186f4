"""GPU tests of the drop-in driver, written like the reference's own
hicexplorer/test/general/test_hicCorrectMatrix.py (argv list -> main() -> compare the output file with the golden
content).  Inputs are rebuilt from the committed fixtures (tests/golden/*.npz) because /root/reference does not
exist on the GPU box."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200 import hdf5lite
from hicexplorer_b200.hicCorrectMatrix import main
from hicexplorer_b200.hicmatrix_lite import hiCMatrix
from oracle import ice_oracle, kr_oracle
from tests.helpers import csr_from_npz, full_from_pixels

pytestmark = pytest.mark.gpu


def _small_input(tmp_path, z, pad_chrom=True):
    """small_test_matrix restricted to chrUextra+chr3LHet (+ an unrelated chromosome that --chromosomes drops)."""
    ma = hiCMatrix()
    m = csr_from_npz(z, "in")
    iv = [(c.decode(), int(s), int(e), float(x)) for c, s, e, x in zip(z["chr"], z["start"], z["end"], z["extra"])]
    if pad_chrom:
        k = 40
        rng = np.random.default_rng(0)
        blk = sparse.random(k, k, 0.2, random_state=rng, format="csr")
        blk = sparse.csr_matrix(np.round((blk + blk.T) * 10))
        m = sparse.block_diag([blk, m], format="csr")
        iv = [("chrOther", i * 5000, (i + 1) * 5000, 1.0) for i in range(k)] + iv
    ma.matrix = sparse.csr_matrix(m)
    ma.cut_intervals = iv
    ma._index_intervals()
    path = str(tmp_path / "small_test_matrix.h5")
    ma.save(path)
    return path


def _read_h5(path):
    f = hdf5lite.File(path)
    shape = tuple(int(v) for v in f["matrix/shape"].read())
    m = sparse.csr_matrix((f["matrix/data"].read(), f["matrix/indices"].read(), f["matrix/indptr"].read()), shape=shape)
    return f, m


def test_correct_matrix_ICE(tmp_path, small_ice):
    """mirrors test_correct_matrix_ICE (test_hicCorrectMatrix.py:29-49)"""
    z = small_ice
    src = _small_input(tmp_path, z)
    out = str(tmp_path / "corrected.h5")
    bed = str(tmp_path / "filtered.bed")
    main("correct --matrix {} --correctionMethod ICE --chromosomes chrUextra chr3LHet --iterNum 500 "
         "--outFileName {} --filteredBed {} --filterThreshold -1.5 5.0".format(src, out, bed).split())
    f, m = _read_h5(out)
    assert np.array_equal(m.indices, z["golden_indices"]) and np.array_equal(m.indptr, z["golden_indptr"])
    np.testing.assert_allclose(m.data, z["golden_data"], rtol=1e-8)        # reference: assert_equal
    assert np.array_equal(f["nan_bins"].read(), z["golden_nan_bins"])
    cf = f["correction_factors"].read()
    assert cf.shape == (6313,)
    assert np.array_equal(np.isnan(cf), np.isnan(z["golden_correction_factors"]))
    ok = ~np.isnan(cf)
    np.testing.assert_allclose(cf[ok], z["golden_correction_factors"][ok], rtol=1e-9)
    assert [c.decode() for c in f["intervals/chr_list"].read()[[0, -1]]] == ["chrUextra", "chr3LHet"]
    assert np.array_equal(f["intervals/start_list"].read(), z["start"]) and np.array_equal(f["intervals/end_list"].read(), z["end"])
    got = sorted(open(bed).read().splitlines())
    want = sorted(str(z["filtered_bed"]).splitlines())
    assert len(got) == len(want) == 64
    for a, b in zip(got, want):                       # the reference tolerates 1 differing char per line (:19-26)
        assert a.split("\t")[:3] == b.split("\t")[:3] and len(a.split("\t")) == 6


def test_correct_matrix_KR_H5(tmp_path, small_ice, small_kr):
    """mirrors test_correct_matrix_KR_H5 (:52-69): data to 5 decimals"""
    src = _small_input(tmp_path, small_ice)
    out = str(tmp_path / "kr.h5")
    main("correct --matrix {} --correctionMethod KR --chromosomes chrUextra chr3LHet --outFileName {}".format(src, out).split())
    f, m = _read_h5(out)
    assert np.array_equal(m.indices, small_kr["golden_indices"]) and np.array_equal(m.indptr, small_kr["golden_indptr"])
    np.testing.assert_almost_equal(m.data, small_kr["golden_data"], decimal=5)
    cf = f["correction_factors"].read()
    assert cf.shape == (6313, 1)                                                # (n,1) for KR
    ref = kr_oracle.kr_pipeline(csr_from_npz(small_ice, "in"), rescale=True)
    np.testing.assert_allclose(cf.ravel(), ref["x_out"], rtol=5e-6)


def _gm_cool(tmp_path, g):
    ma = hiCMatrix()
    ma.matrix = full_from_pixels(g)
    off = g["chrom_offset"]
    ma.cut_intervals = [(str(c + 1), int((i - off[c]) * 2500000), int((i - off[c] + 1) * 2500000), 1.0)
                        for c in range(len(off) - 1) for i in range(off[c], off[c + 1])]
    ma._index_intervals()
    path = str(tmp_path / "gm12878_raw_values.cool")
    ma.save(path)
    return path


def test_correct_matrix_KR_cool(tmp_path, gm12878):
    """mirrors test_correct_matrix_KR_cool (:72-91): .cool keeps the raw counts, weights = rescaled x"""
    src = _gm_cool(tmp_path, gm12878)
    out = str(tmp_path / "kr.cool")
    main("correct --matrix {} --correctionMethod KR --outFileName {}".format(src, out).split())
    new = hiCMatrix(out)
    raw = full_from_pixels(gm12878)
    assert abs(new.matrix - raw).sum() == 0
    assert new.matrix.sum() // 2 > 3e9                               # lower half of the reference's sanity bound
    ref = kr_oracle.kr_pipeline(raw, rescale=True)
    np.testing.assert_allclose(np.asarray(new.correction_factors).ravel(), ref["x_out"], rtol=1e-9)


def test_correct_matrix_KR_partial_cool(tmp_path, gm12878, kr_partial):
    """mirrors test_correct_matrix_KR_partial_cool (:94-109): single chromosome through pChrnameList"""
    src = _gm_cool(tmp_path, gm12878)
    out = str(tmp_path / "kr_partial.cool")
    main("correct --matrix {} --correctionMethod KR --chromosomes 3 --outFileName {}".format(src, out).split())
    new = hiCMatrix(out)
    assert new.matrix.shape == (80, 80)
    w = np.asarray(new.correction_factors).ravel()
    ratio = kr_partial["weight"] / w
    assert np.ptp(ratio) < 1e-9 and abs(ratio.mean() - 1) < 5e-6      # the shipped file, up to its rescale constant


def test_correct_matrix_ICE_cool_and_perchr(tmp_path, gm12878):
    g = gm12878
    src = _gm_cool(tmp_path, g)
    out = str(tmp_path / "ice.cool")
    main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 --outFileName {}".format(src, out).split())
    new = hiCMatrix(out)
    ref = ice_oracle.correct_ice_pipeline(full_from_pixels(g), -1.5, 5.0, max_iter=500)
    cf = np.asarray(new.correction_factors).ravel()
    np.testing.assert_array_equal(np.flatnonzero(np.isnan(cf)), ref["nan_bins"])
    np.testing.assert_allclose(cf[ref["kept"]], ref["ice"]["bias"], rtol=1e-9)
    np.testing.assert_allclose(sparse.triu(new.matrix).tocsr().data, ref["upper"].data, rtol=1e-8)
    # --perchr: per-chromosome filter + independent corrections, trans contacts dropped in the .h5 output
    outp = str(tmp_path / "ice_perchr.h5")
    main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 --perchr --iterNum 300 "
         "--outFileName {}".format(src, outp).split())
    f, m = _read_h5(outp)
    full = full_from_pixels(g)
    keep1 = np.setdiff1d(np.arange(full.shape[0]), g["zero_bins"])
    kept = keep1[np.setdiff1d(np.arange(len(keep1)), g["outliers_rel_perchr"])]
    cf = f["correction_factors"].read()
    assert np.array_equal(np.flatnonzero(~np.isnan(cf)), kept)
    off = g["chrom_offset"]
    chrom_of = np.searchsorted(off, np.arange(full.shape[0]), side="right") - 1
    coo = m.tocoo()
    assert np.all(chrom_of[coo.row] == chrom_of[coo.col])               # no inter-chromosomal entries survive
    c = 2                                                                # check one chromosome against the oracle
    ids = kept[chrom_of[kept] == c]
    blk = ice_oracle.scrub(ice_oracle.select_bins(full, ids))
    refb = ice_oracle.ice(blk, max_iter=300)
    np.testing.assert_allclose(cf[ids], refb["bias"], rtol=1e-9)


def test_skip_diagonal_and_sequenced_count_cutoff(tmp_path, gm12878):
    """--skipDiagonal (hicCorrectMatrix.py:648-649) and --sequencedCountCutoff (:675-686) against the oracle run with
    the same sequence of host operations the reference performs."""
    g = gm12878
    full = full_from_pixels(g)
    n = full.shape[0]
    rng = np.random.default_rng(7)
    coverage = rng.uniform(0.3, 1.0, n)
    ma = hiCMatrix()
    ma.matrix = full
    off = g["chrom_offset"]
    ma.cut_intervals = [(str(c + 1), int((i - off[c]) * 2500000), int((i - off[c] + 1) * 2500000), float(coverage[i]))
                        for c in range(len(off) - 1) for i in range(off[c], off[c + 1])]
    ma._index_intervals()
    src = str(tmp_path / "in.h5")
    ma.save(src)
    out = str(tmp_path / "out.h5")
    main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 --skipDiagonal --sequencedCountCutoff 0.5 "
         "--iterNum 300 --outFileName {}".format(src, out).split())
    f, m = _read_h5(out)
    # oracle: zero bins -> scrub -> diagonal removed -> MAD mask -> coverage mask -> ICE
    keep1 = np.setdiff1d(np.arange(n), ice_oracle.zero_bins(full))
    st1 = ice_oracle.scrub(ice_oracle.select_bins(full, keep1)).tolil()
    st1.setdiag(0)
    st1 = st1.tocsr()
    st1.eliminate_zeros()
    outl = ice_oracle.outlier_bins(st1, -1.5, 5.0)
    keep2 = np.delete(np.arange(st1.shape[0]), outl)
    low_cov = coverage[keep1[keep2]] < 0.5
    kept = keep1[keep2][~low_cov]
    ref = ice_oracle.ice(ice_oracle.select_bins(st1, keep2[~low_cov]), max_iter=300)
    cf = f["correction_factors"].read()
    assert np.array_equal(np.flatnonzero(~np.isnan(cf)), kept)
    np.testing.assert_allclose(cf[kept], ref["bias"], rtol=1e-9)
    assert m.diagonal().sum() == 0                                   # the diagonal is gone from the output
    up = sparse.triu(ref["matrix"], 0, format="csr")
    up.eliminate_zeros()
    np.testing.assert_allclose(np.sort(m.data), np.sort(up.data), rtol=1e-8)


def test_trans_and_inflation_cutoff(tmp_path, gm12878):
    """--transCutoff / --inflationCutoff (hicCorrectMatrix.py:695-699, 765-775): inflated bins become nan_bins with NaN
    factors and leave the output matrix; --inflationCutoff alone exits 1 (the reference dies with a NameError)"""
    g = gm12878
    src = _gm_cool(tmp_path, g)
    full = full_from_pixels(g)
    out = str(tmp_path / "ice_infl.h5")
    main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 --transCutoff 0.05 --inflationCutoff 1.7 "
         "--outFileName {}".format(src, out).split())
    ref = ice_oracle.correct_ice_pipeline(full, -1.5, 5.0, max_iter=500, chr_offsets=g["chrom_offset"], trans_cutoff=0.05,
                                          inflation_cutoff=1.7)
    assert 0 < len(ref["inflated_rel"]) < 1000
    f, m = _read_h5(out)
    assert np.array_equal(f["nan_bins"].read(), ref["nan_bins"])
    assert np.array_equal(m.indptr, ref["upper"].indptr) and np.array_equal(m.indices, ref["upper"].indices)
    np.testing.assert_allclose(m.data, ref["upper"].data, rtol=1e-8)
    cf = f["correction_factors"].read()
    assert np.array_equal(np.isnan(cf), np.isnan(ref["correction_factors"]))
    np.testing.assert_allclose(cf[ref["kept"]], ref["correction_factors"][ref["kept"]], rtol=1e-9)
    with pytest.raises(SystemExit) as exc:
        main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 --inflationCutoff 3 "
             "--outFileName {}".format(src, out).split())
    assert exc.value.code == 1


def test_kr_from_the_files_upper_triangle(tmp_path, gm12878):
    """the default method on an untouched input: upper triangle uploaded as stored, symmetric store built on the device,
    balanced matrix back in file layout -- genome-wide and --perchr, .h5 (matrix + factors) and .cool (factors only)"""
    g = gm12878
    src = _gm_cool(tmp_path, g)
    raw = full_from_pixels(g)
    ref = kr_oracle.kr_pipeline(raw, rescale=True)
    out = str(tmp_path / "kr.h5")
    main("correct --matrix {} --correctionMethod KR --outFileName {}".format(src, out).split())
    f, m = _read_h5(out)
    assert np.array_equal(m.indices, ref["upper"].indices) and np.array_equal(m.indptr, ref["upper"].indptr)
    np.testing.assert_allclose(m.data, ref["upper"].data, rtol=1e-8)
    cf = f["correction_factors"].read()
    assert cf.shape == (raw.shape[0], 1)
    np.testing.assert_allclose(cf.ravel(), ref["x_out"], rtol=1e-9)
    # --perchr: every chromosome balanced on its own, trans entries dropped from the .h5 matrix
    offs = g["chrom_offset"]
    outp = str(tmp_path / "kr_perchr.h5")
    main("correct --matrix {} --correctionMethod KR --perchr --outFileName {}".format(src, outp).split())
    fp, mp = _read_h5(outp)
    cfp = fp["correction_factors"].read().ravel()
    for c in (0, 5, len(offs) - 2):
        lo, hi = int(offs[c]), int(offs[c + 1])
        blk = kr_oracle.kr_pipeline(sparse.csr_matrix(raw[lo:hi, lo:hi]), rescale=True)
        np.testing.assert_allclose(cfp[lo:hi], blk["x_out"], rtol=1e-9)
        got = mp[lo:hi, lo:hi].tocsr()
        got.sort_indices()
        assert np.array_equal(got.indices, blk["upper"].indices)
        np.testing.assert_allclose(got.data, blk["upper"].data, rtol=1e-8)
    assert mp[int(offs[0]):int(offs[1]), int(offs[1]):].nnz == 0          # no trans entries
    outc = str(tmp_path / "kr_perchr.cool")
    main("correct --matrix {} --correctionMethod KR --perchr --outFileName {}".format(src, outc).split())
    newc = hiCMatrix(outc)
    assert abs(newc.matrix - raw).sum() == 0                               # .cool keeps the raw counts
    lo, hi = int(offs[2]), int(offs[3])
    unres = kr_oracle.kr_pipeline(sparse.csr_matrix(raw[lo:hi, lo:hi]), rescale=False)
    np.testing.assert_allclose(np.asarray(newc.correction_factors).ravel()[lo:hi], unres["x_out"], rtol=1e-9)


def test_chromosome_subset_in_file_order_is_selected_on_the_device(tmp_path, gm12878):
    """--chromosomes with a subset in file order: the upper triangle is uploaded untouched and the other bins are
    deleted on the device (hicmatrix reorderChromosomes on the host otherwise); ICE and KR"""
    g = gm12878
    src = _gm_cool(tmp_path, g)
    offs = g["chrom_offset"]
    chroms = [2, 3, 7]                                        # names "2", "3", "7" of _gm_cool: file order
    keep = np.concatenate([np.arange(offs[c - 1], offs[c]) for c in chroms])
    sub = sparse.csr_matrix(full_from_pixels(g)[keep, :][:, keep])
    sub.sort_indices()
    ma = hiCMatrix(src)
    ma.reorderChromosomes([str(c) for c in chroms])
    assert ma.upper_for_device() is not None and ma.upper_drop_mask().sum() == g["nbins"] - len(keep)
    assert (ma.matrix != sub).nnz == 0 and ma.upper_for_device() is None          # host view applies the selection
    out = str(tmp_path / "ice_sub.h5")
    main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 --chromosomes 2 3 7 "
         "--outFileName {}".format(src, out).split())
    ref = ice_oracle.correct_ice_pipeline(sub, -1.5, 5.0, max_iter=500)
    f, m = _read_h5(out)
    assert m.shape == sub.shape
    assert np.array_equal(f["nan_bins"].read(), ref["nan_bins"])
    assert np.array_equal(m.indptr, ref["upper"].indptr) and np.array_equal(m.indices, ref["upper"].indices)
    np.testing.assert_allclose(m.data, ref["upper"].data, rtol=1e-8)
    assert [c.decode() for c in np.unique(f["intervals/chr_list"].read())] == ["2", "3", "7"]
    outk = str(tmp_path / "kr_sub.h5")
    main("correct --matrix {} --correctionMethod KR --chromosomes 2 3 7 --outFileName {}".format(src, outk).split())
    kref = kr_oracle.kr_pipeline(sub, rescale=True)
    fk, mk = _read_h5(outk)
    np.testing.assert_allclose(fk["correction_factors"].read().ravel(), kref["x_out"], rtol=1e-9)
    assert np.array_equal(mk.indices, kref["upper"].indices)
    np.testing.assert_allclose(mk.data, kref["upper"].data, rtol=1e-8)
    outc = str(tmp_path / "kr_sub.cool")
    main("correct --matrix {} --correctionMethod KR --chromosomes 2 3 7 --outFileName {}".format(src, outc).split())
    newc = hiCMatrix(outc)
    assert abs(newc.matrix - sub).sum() == 0
    np.testing.assert_allclose(np.asarray(newc.correction_factors).ravel(), kref["x_out"], rtol=1e-9)


def test_device_side_chromosome_selection_equals_a_presliced_input(tmp_path, gm12878):
    """--chromosomes (subset in file order, selected on the device) gives byte-for-byte the matrix and factors of the
    same command run on an input that holds only those chromosomes -- also with --perchr"""
    g = gm12878
    src = _gm_cool(tmp_path, g)
    pre = hiCMatrix(src)
    pre.reorderChromosomes(["4", "9", "10"])
    _ = pre.matrix                                   # host-side selection
    sliced = str(tmp_path / "sliced.cool")
    pre.save(sliced)
    for extra in ("", "--perchr"):
        a, b = str(tmp_path / "a.h5"), str(tmp_path / "b.h5")
        main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 --chromosomes 4 9 10 {} "
             "--outFileName {}".format(src, extra, a).split())
        main("correct --matrix {} --correctionMethod ICE --filterThreshold -1.5 5 {} --outFileName {}".format(
            sliced, extra, b).split())
        fa, ma_ = _read_h5(a)
        fb, mb_ = _read_h5(b)
        assert np.array_equal(ma_.indptr, mb_.indptr) and np.array_equal(ma_.indices, mb_.indices)
        np.testing.assert_array_equal(ma_.data, mb_.data)
        np.testing.assert_array_equal(fa["correction_factors"].read(), fb["correction_factors"].read())
        assert np.array_equal(fa["nan_bins"].read(), fb["nan_bins"].read())


def test_diagnostic_plot_statistics(tmp_path, gm12878):
    """`hicCorrectMatrix diagnostic_plot` (hicCorrectMatrix.py:425-545): the histogram, the first local minimum and its
    modified z-score ("mad threshold") equal the oracle's restatement; whole matrix and --perchr; the figure is written"""
    import os
    g = gm12878
    src = _gm_cool(tmp_path, g)
    full = full_from_pixels(g)
    out = str(tmp_path / "diag.svg")
    panels = main("diagnostic_plot --matrix {} --plotName {}".format(src, out).split())
    ref = ice_oracle.diagnostic_histograms(full)
    assert len(panels) == 1 and os.path.getsize(out) > 1000
    np.testing.assert_array_equal(panels[0]["values"], ref[0]["values"])
    np.testing.assert_array_equal(panels[0]["counts"], ref[0]["counts"])
    assert panels[0]["threshold"] == ref[0]["threshold"] and panels[0]["mad_threshold"] == ref[0]["mad_threshold"]
    cut = float(np.percentile(ref[0]["values"], 80))
    px = main("diagnostic_plot --matrix {} --plotName {} --xMax {}".format(src, out, cut).split())
    rx = ice_oracle.diagnostic_histograms(full, x_max=cut)
    np.testing.assert_array_equal(px[0]["counts"], rx[0]["counts"])
    assert px[0]["mad_threshold"] == rx[0]["mad_threshold"]
    offs = g["chrom_offset"]
    ranges = list(zip(offs[:-1], offs[1:]))
    outp = str(tmp_path / "diag_perchr.svg")
    pan = main("diagnostic_plot --matrix {} --plotName {} --perchr".format(src, outp).split())
    refp = ice_oracle.diagnostic_histograms(full, chr_ranges=ranges)
    assert sum(p["mad_threshold"] is not None for p in refp) > 10
    assert len(pan) == len(ranges) and os.path.getsize(outp) > 1000
    for a, b in zip(pan, refp):
        np.testing.assert_array_equal(a["values"], b["values"])
        assert (a["counts"] is None) == (b["counts"] is None)
        if a["counts"] is not None:
            np.testing.assert_array_equal(a["counts"], b["counts"])
        assert a["mad_threshold"] == b["mad_threshold"]
