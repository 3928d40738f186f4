"""CPU tests of the host side above the C ABI: file formats (hdf5lite / hicmatrix_lite), the CLI surface."""
import numpy as np
import pytest
from scipy import sparse

from hicexplorer_b200 import hdf5lite
from hicexplorer_b200.hicCorrectMatrix import parse_arguments
from hicexplorer_b200.hicmatrix_lite import hiCMatrix
from oracle import ref_import
from tests.helpers import csr_from_npz, full_from_pixels


def _ma_from_small(small_ice):
    ma = hiCMatrix()
    ma.matrix = csr_from_npz(small_ice, "in")
    ma.cut_intervals = [(c.decode(), int(s), int(e), float(x)) for c, s, e, x in
                        zip(small_ice["chr"], small_ice["start"], small_ice["end"], small_ice["extra"])]
    ma._index_intervals()
    return ma


def test_h5_and_cool_round_trip(tmp_path, small_ice):
    ma = _ma_from_small(small_ice)
    assert ma.getChrNames() == ["chrUextra", "chr3LHet"] and list(ma.chromosome_offsets()) == [0, 5801, 6313]
    ref_full = ma.matrix.copy()
    ma.maskBins(small_ice["zero_bins"])
    assert ma.matrix.shape == (183, 183)
    ma.maskBins(small_ice["outliers_rel"])                      # second call: union of both masks
    assert ma.matrix.shape == (119, 119)
    ma.setCorrectionFactors(np.arange(119, dtype=np.float64))
    for ext in ("h5", "cool"):
        m2 = _ma_from_small(small_ice)
        m2.maskBins(small_ice["zero_bins"])
        m2.maskBins(small_ice["outliers_rel"])
        m2.setCorrectionFactors(np.arange(119, dtype=np.float64))
        path = str(tmp_path / ("out." + ext))
        m2.save(path)
        back = hiCMatrix(path)
        assert back.matrix.shape == (6313, 6313)
        assert abs(back.matrix - sparse.csr_matrix(ref_full.multiply(back.matrix != 0))).sum() == 0
        cf = np.asarray(back.correction_factors).ravel()
        assert np.array_equal(np.flatnonzero(np.isnan(cf)), small_ice["golden_nan_bins"])
        assert np.array_equal(cf[~np.isnan(cf)], np.arange(119, dtype=np.float64))
        assert [iv[:3] for iv in back.cut_intervals[:2]] == [("chrUextra", 0, 5000), ("chrUextra", 5000, 10000)]
    f = hdf5lite.File(str(tmp_path / "out.h5"))
    assert f.attrs["TITLE"] == "HiCExplorer matrix"
    assert sorted(f.keys()) == ["correction_factors", "intervals", "matrix", "nan_bins"]
    assert f["matrix/indices"].dtype == np.int32 and f["matrix/indptr"].dtype == np.int32
    assert f["matrix/data"].dtype == np.float64 and f["nan_bins"].dtype == np.int64
    assert np.array_equal(f["nan_bins"].read(), small_ice["golden_nan_bins"])
    up = sparse.csr_matrix((f["matrix/data"].read(), f["matrix/indices"].read(), f["matrix/indptr"].read()),
                           shape=(6313, 6313))
    assert sparse.tril(up, -1).nnz == 0
    c = hdf5lite.File(str(tmp_path / "out.cool"))
    assert c.attrs["format"] == "HDF5::Cooler" and c.attrs["nbins"] == 6313
    assert c["bins/chrom"].enum == {"chrUextra": 0, "chr3LHet": 1}
    b1, b2 = c["pixels/bin1_id"].read(), c["pixels/bin2_id"].read()
    assert np.all(b1 <= b2) and np.all(np.diff(b1) >= 0)
    assert np.array_equal(c["indexes/bin1_offset"].read(), np.concatenate([[0], np.cumsum(np.bincount(b1, minlength=6313))]))


def test_cool_with_many_chromosomes_round_trip(tmp_path, gm12878):
    ma = hiCMatrix()
    ma.matrix = full_from_pixels(gm12878)
    off = gm12878["chrom_offset"]
    ma.cut_intervals = [("chr%d" % c, int((i - off[c]) * 2500000), int((i - off[c] + 1) * 2500000), 1.0)
                        for c in range(len(off) - 1) for i in range(off[c], off[c + 1])]
    ma._index_intervals()
    path = str(tmp_path / "g.cool")
    ma.save(path)
    back = hiCMatrix(path)
    assert abs(back.matrix - ma.matrix).sum() == 0 and len(back.getChrNames()) == 25
    one = hiCMatrix(path, pChrnameList=["chr2"])
    a, b = int(off[2]), int(off[3])
    assert abs(one.matrix - ma.matrix[a:b, a:b]).sum() == 0


def test_cli_defaults_mirror_the_reference():
    a = parse_arguments().parse_args(["correct", "-m", "x.h5", "-o", "y.h5"])
    assert a.correctionMethod == "KR" and a.iterNum == 500 and a.filterThreshold is None
    assert a.perchr is False and a.skipDiagonal is False and a.chromosomes is None and a.filteredBed is None
    assert a.inflationCutoff is None and a.transCutoff is None and a.sequencedCountCutoff is None
    assert a.tolerance == 1e-5 and a.device == 0                      # additive flags
    b = parse_arguments().parse_args(["correct", "--matrix", "x.h5", "--outFileName", "y.h5", "--correctionMethod", "ICE",
                                      "-t", "-1.5", "5", "-n", "20", "--chromosomes", "chr1", "chr2", "--perchr", "-s",
                                      "--filteredBed", "f.bed", "--verbose"])
    assert b.filterThreshold == [-1.5, 5.0] and b.iterNum == 20 and b.chromosomes == ["chr1", "chr2"] and b.perchr
    with pytest.raises(SystemExit):
        parse_arguments().parse_args(["correct", "-m", "x.h5"])         # -o is required
    with pytest.raises(SystemExit):
        parse_arguments().parse_args(["correct", "-m", "x", "-o", "y", "--correctionMethod", "XX"])


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_cli_flags_are_a_superset_of_the_reference_parser():
    ref = ref_import.load()[3]()
    mine = parse_arguments()

    def flags(parser):
        sub = [a for a in parser._actions if a.__class__.__name__ == "_SubParsersAction"][0]
        out = {}
        for act in sub.choices["correct"]._actions:
            for s in act.option_strings:
                out[s] = (act.default, act.nargs, act.required)
        return out

    r, m = flags(ref), flags(mine)
    for flag, spec in r.items():
        assert flag in m, "missing flag " + flag
        assert m[flag] == spec, (flag, m[flag], spec)
    assert set(m) - set(r) == {"--tolerance", "--device", "--clipTrans"}


def test_ice_without_thresholds_exits_1_like_the_reference(tmp_path, small_ice):
    from hicexplorer_b200.hicCorrectMatrix import main
    ma = _ma_from_small(small_ice)
    src = str(tmp_path / "in.h5")
    ma.save(src)
    with pytest.raises(SystemExit) as ei:
        main(["correct", "-m", src, "-o", str(tmp_path / "o.h5"), "--correctionMethod", "ICE"])
    assert ei.value.code == 1


@pytest.mark.parametrize("count", [1, 2, 3, 10, 1001, 65537])
@pytest.mark.parametrize("q", [0.0, 100.0, 99.9995, 50.0, 12.5, 99.95])
def test_percentile_position_reproduces_numpy(count, q):
    """the host half of --transCutoff: numpy's 'linear' percentile from two order statistics, bit for bit"""
    from hicexplorer_b200.store import lerp_like_numpy, percentile_position
    rng = np.random.default_rng(count)
    x = np.floor(rng.exponential(5.0, count)) * rng.random(count)
    srt = np.sort(x)
    k, gamma = percentile_position(count, q)
    got = lerp_like_numpy(srt[k], srt[min(k + 1, count - 1)], gamma)
    assert got == np.percentile(x, q)


def test_oracle_trunc_trans_semantics():
    from oracle import ice_oracle
    m = sparse.csr_matrix(np.array([[5, 1, 9, 0], [1, 5, 0, 2], [9, 0, 5, 1], [0, 2, 1, 5]], dtype=float))
    same, mx = ice_oracle.trunc_trans(m, [0, 2, 4], 50.0, clip=False)
    assert mx == np.percentile([9, 2, 9, 2], 50.0) and (same != m).nnz == 0
    clipped, _ = ice_oracle.trunc_trans(m, [0, 2, 4], 50.0, clip=True)
    assert clipped[0, 2] == mx and clipped[2, 0] == mx and clipped[0, 1] == 1 and clipped[1, 3] == 2


def test_mcool_uri_selects_a_resolution(tmp_path):
    """`file.mcool::/resolutions/N` (cooler URI syntax) loads the matrix stored under that group"""
    import numpy as np
    from hicexplorer_b200 import hdf5lite
    ma = hiCMatrix()
    up = sparse.triu(sparse.csr_matrix(np.arange(36, dtype=np.float64).reshape(6, 6) % 5), 0, format="csr")
    ma.matrix = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    ma.cut_intervals = [("chr1" if i < 4 else "chr2", (i % 4) * 100, (i % 4 + 1) * 100, 1.0) for i in range(6)]
    ma._index_intervals()
    single = str(tmp_path / "single.cool")
    ma.save(single)
    src = hdf5lite.File(single)
    w = hdf5lite.Writer({"format": "HDF5::MCOOL"})
    for grp in ("bins", "chroms", "indexes", "pixels"):
        for name in src[grp].keys():
            ds = src[grp + "/" + name]
            w.dataset("/resolutions/100/%s/%s" % (grp, name), ds.read(), enum=ds.enum)
    path = str(tmp_path / "multi.mcool")
    w.save(path)
    got = hiCMatrix(path + "::/resolutions/100")
    assert (got.matrix != ma.matrix).nnz == 0 and got.getChrNames() == ["chr1", "chr2"]


def test_diagnostic_plot_svg_writer(tmp_path):
    """the hand-written SVG of `diagnostic_plot` (used when matplotlib is not installed) is well-formed and holds one
    bar per histogram bin, the threshold line and both axis captions"""
    import xml.etree.ElementTree as ET
    from hicexplorer_b200.hicCorrectMatrix import _draw_svg
    rng = np.random.default_rng(0)
    panels = []
    for title in ("chr1", "chr2", None):
        values = rng.gamma(5.0, 100.0, size=3000)
        counts, edges = np.histogram(values, 100)
        panels.append(dict(title=title, values=values, counts=counts, edges=edges, threshold=float(edges[7]),
                           mad_threshold=-1.8, median=float(np.median(values)), mad=80.0))
    panels.append(dict(title="empty", values=np.zeros(0), counts=None, edges=None, threshold=None, mad_threshold=None,
                       median=0.0, mad=0.0))
    path = str(tmp_path / "plot.svg")
    _draw_svg(panels, path)
    root = ET.parse(path).getroot()
    ns = "{http://www.w3.org/2000/svg}"
    bars = [r for r in root.iter(ns + "rect") if r.get("fill") == "green"]
    assert len(bars) == 300
    assert len(list(root.iter(ns + "line"))) == 3
    texts = [t.text for t in root.iter(ns + "text")]
    assert texts.count("total counts per bin") == 4 and texts.count("modified z-score") == 4 and "chr2" in texts
    other = str(tmp_path / "plot.png")
    _draw_svg(panels[:1], other)                         # not an .svg name: written next to it, with a warning
    assert (tmp_path / "plot.png.svg").exists()


def test_keep_only_these_chr_keeps_the_file_order(tmp_path, gm12878):
    """hicTransform --chromosomes goes through hicmatrix keepOnlyTheseChr (hicTransform.py:156): a selection in FILE
    order, unlike reorderChromosomes which follows the order of the list"""
    from hicexplorer_b200.hicmatrix_lite import hiCMatrix
    from tests.helpers import full_from_pixels
    ma = hiCMatrix()
    ma.matrix = full_from_pixels(gm12878)
    off = gm12878["chrom_offset"]
    ma.cut_intervals = [(str(c + 1), int((i - off[c]) * 100), int((i - off[c] + 1) * 100), 1.0)
                        for c in range(len(off) - 1) for i in range(off[c], off[c + 1])]
    ma._index_intervals()
    path = str(tmp_path / "m.cool")
    ma.save(path)
    a = hiCMatrix(path)
    a.keepOnlyTheseChr(["7", "2", "3"])
    assert a.getChrNames() == ["2", "3", "7"]
    b = hiCMatrix(path)
    b.reorderChromosomes(["7", "2", "3"])
    assert b.getChrNames() == ["7", "2", "3"]
    assert a.matrix.shape == b.matrix.shape and a.matrix.sum() == b.matrix.sum()
    lo, hi = a.getChrBinRange("2")
    want = full_from_pixels(gm12878)[off[1]:off[2], off[1]:off[2]]
    assert (a.matrix[lo:hi, lo:hi] != want).nnz == 0
