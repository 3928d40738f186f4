"""Generate the committed golden vectors under ``tests/golden/``.

Runs ONLY in the build container (needs ``/root/reference``): it reads the
reference's own fixtures/golden files with ``hicexplorer_b200.hdf5lite`` and
runs the reference's *own* ``iterativeCorrection`` / ``MAD`` /
``filter_by_zscore`` (imported unmodified through ``oracle.ref_import``) to
produce expected outputs.  The resulting ``.npz`` files travel to the GPU box;
``/root/reference`` does not.

    python tests/golden/make_golden.py

Files written:
  small_ice.npz   config 1 of BASELINE.json: small_test_matrix.h5 restricted to
                  chrUextra + chr3LHet, with every stage of the ICE pipeline and
                  the content of the reference's golden output file.
  small_kr.npz    same input, the reference's golden KR output file.
  gm12878.npz     gm12878_raw_values.cool pixels + reference ICE result
                  (mask, bias, iterations, sampled corrected values).
  kr_partial.npz  kr_partial.cool (chromosome 3 of gm12878): raw counts + weights.
"""
import os
import sys

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from hicexplorer_b200 import hdf5lite  # noqa: E402
from oracle import ref_import  # noqa: E402

TD = "/root/reference/hicexplorer/test/test_data/"


class _Ma(object):
    """Object with the two attributes filter_by_zscore touches."""

    def __init__(self, matrix):
        self.matrix = matrix


def read_h5_matrix(path):
    f = hdf5lite.File(path)
    shape = tuple(int(v) for v in f["matrix/shape"].read())
    m = sparse.csr_matrix((f["matrix/data"].read(), f["matrix/indices"].read(), f["matrix/indptr"].read()),
                          shape=shape)
    iv = dict(chr=f["intervals/chr_list"].read(), start=f["intervals/start_list"].read(),
              end=f["intervals/end_list"].read(), extra=f["intervals/extra_list"].read())
    extra = {}
    if "nan_bins" in f:
        extra["nan_bins"] = f["nan_bins"].read()
    if "correction_factors" in f:
        extra["correction_factors"] = f["correction_factors"].read()
    return m, iv, extra


def read_cool(path):
    f = hdf5lite.File(path)
    b1 = f["pixels/bin1_id"].read()
    b2 = f["pixels/bin2_id"].read()
    cnt = f["pixels/count"].read()
    nb = int(f["bins/start"].shape[0])
    out = dict(bin1=b1, bin2=b2, count=cnt, nbins=nb, chrom_offset=f["indexes/chrom_offset"].read(),
               start=f["bins/start"].read(), end=f["bins/end"].read(), chrom_id=f["bins/chrom"].read(),
               chrom_names=f["chroms/name"].read())
    if "weight" in f["bins"].keys():
        out["weight"] = f["bins/weight"].read()
    return out


def full_from_pixels(b1, b2, cnt, n):
    up = sparse.csr_matrix((cnt, (b1, b2)), shape=(n, n))
    return sparse.csr_matrix(up + sparse.triu(up, 1).T)


def main():
    iterativeCorrection, MAD, filter_by_zscore, _ = ref_import.load()

    # ---- config 1: small_test_matrix.h5, chrUextra + chr3LHet ------------------
    m, iv, _ = read_h5_matrix(TD + "small_test_matrix.h5")
    full = sparse.csr_matrix(m + sparse.triu(m, 1).T)
    order = np.concatenate([np.flatnonzero(iv["chr"] == c) for c in (b"chrUextra", b"chr3LHet")])
    sub = sparse.csr_matrix(full[order, :][:, order])
    sub.sort_indices()
    n = sub.shape[0]
    rs = np.asarray(sub.sum(axis=1)).flatten()
    zero = np.flatnonzero(rs == 0)
    keep1 = np.flatnonzero(rs != 0)
    st1 = sparse.csr_matrix(sub[keep1, :][:, keep1]).astype(np.float64)
    outl = np.array(filter_by_zscore(_Ma(st1), -1.5, 5.0), dtype=np.int64)
    mad = MAD(np.asarray(st1.sum(axis=1)).flatten() - st1.diagonal())
    keep2 = np.delete(np.arange(st1.shape[0]), outl)
    ice_in = sparse.csr_matrix(st1[keep2, :][:, keep2])
    ice_in.sort_indices()
    W, bias = iterativeCorrection(ice_in.copy(), M=500)
    W.sort_indices()
    gm, giv, gextra = read_h5_matrix(TD + "hicCorrectMatrix/small_test_matrix_ICEcorrected_chrUextra_chr3LHet.h5")
    bed = open(TD + "hicCorrectMatrix/filtered.bed").read()
    np.savez_compressed(
        os.path.join(HERE, "small_ice.npz"),
        in_indptr=sub.indptr.astype(np.int64), in_indices=sub.indices.astype(np.int32),
        in_data=sub.data.astype(np.float64), n=np.int64(n),
        chr=iv["chr"][order], start=iv["start"][order], end=iv["end"][order], extra=iv["extra"][order],
        zero_bins=zero.astype(np.int64), outliers_rel=outl, zscore=mad.get_motified_zscores(),
        mad_median=np.float64(mad.median), mad_mad=np.float64(mad.med_abs_deviation),
        ice_in_indptr=ice_in.indptr.astype(np.int64), ice_in_indices=ice_in.indices.astype(np.int32),
        ice_in_data=ice_in.data, ice_out_data=W.data, ice_out_indices=W.indices.astype(np.int32),
        ice_out_indptr=W.indptr.astype(np.int64), ice_bias=bias,
        golden_data=gm.data, golden_indices=gm.indices, golden_indptr=gm.indptr,
        golden_correction_factors=gextra["correction_factors"], golden_nan_bins=gextra["nan_bins"],
        filtered_bed=np.array(bed))

    # ---- KR golden for the same input -----------------------------------------
    km, _kiv, kextra = read_h5_matrix(TD + "hicCorrectMatrix/small_test_matrix_KRcorrected_chrUextra_chr3LHet.h5")
    np.savez_compressed(
        os.path.join(HERE, "small_kr.npz"),
        golden_data=km.data, golden_indices=km.indices, golden_indptr=km.indptr,
        golden_correction_factors=kextra["correction_factors"])

    # ---- gm12878 2.5 Mb: realistic CPU-sized ICE run ----------------------------
    c = read_cool(TD + "hicCorrectMatrix/gm12878_raw_values.cool")
    n = c["nbins"]
    fullg = full_from_pixels(c["bin1"], c["bin2"], c["count"].astype(np.float64), n)
    rs = np.asarray(fullg.sum(axis=1)).flatten()
    zero = np.flatnonzero(rs == 0)
    keep1 = np.flatnonzero(rs != 0)
    st1 = sparse.csr_matrix(fullg[keep1, :][:, keep1]).astype(np.float64)
    outl = np.array(filter_by_zscore(_Ma(st1), -1.5, 5.0), dtype=np.int64)
    mad = MAD(np.asarray(st1.sum(axis=1)).flatten() - st1.diagonal())
    # --perchr variant of the filter on the same matrix (chromosome blocks after zero-bin removal)
    chrom_of_bin = c["chrom_id"][keep1]
    bounds = np.flatnonzero(np.diff(chrom_of_bin)) + 1
    chr_lo = np.concatenate([[0], bounds])
    chr_hi = np.concatenate([bounds, [len(chrom_of_bin)]])

    class _MaChr(_Ma):
        def __init__(self, matrix):
            _Ma.__init__(self, matrix)
            self.interval_trees = dict((i, None) for i in range(len(chr_lo)))

        def getChrBinRange(self, k):
            return int(chr_lo[k]), int(chr_hi[k])

    outl_perchr = np.array(filter_by_zscore(_MaChr(st1.copy()), -1.5, 5.0, perchr=True), dtype=np.int64)
    keep2 = np.delete(np.arange(st1.shape[0]), outl)
    ice_in = sparse.csr_matrix(st1[keep2, :][:, keep2])
    ice_in.sort_indices()
    import logging
    import io
    buf = io.StringIO()
    h = logging.StreamHandler(buf)
    lg = logging.getLogger("hicexplorer.iterativeCorrection")
    lg.addHandler(h)
    lg.setLevel(logging.INFO)
    W, bias = iterativeCorrection(ice_in.copy(), M=500)
    lg.removeHandler(h)
    logged = [ln for ln in buf.getvalue().splitlines() if "iterations used" in ln]
    iters = int(logged[0].split("]")[1].split()[0]) - 1
    W.sort_indices()
    samp = np.arange(0, W.nnz, 997)
    np.savez_compressed(
        os.path.join(HERE, "gm12878.npz"),
        bin1=c["bin1"], bin2=c["bin2"], count=c["count"], nbins=np.int64(n), chrom_offset=c["chrom_offset"],
        zero_bins=zero.astype(np.int64), outliers_rel=outl, outliers_rel_perchr=outl_perchr,
        chr_lo=chr_lo.astype(np.int64), chr_hi=chr_hi.astype(np.int64),
        zscore=mad.get_motified_zscores(), ice_bias=bias, ice_iterations=np.int64(iters),
        ice_sample_idx=samp, ice_sample_val=W.data[samp], ice_rowsum=np.asarray(W.sum(axis=1)).flatten(),
        ice_data_sum=np.float64(W.data.sum()))

    # ---- kr_partial.cool ----------------------------------------------------------
    k = read_cool(TD + "hicCorrectMatrix/kr_partial.cool")
    np.savez_compressed(os.path.join(HERE, "kr_partial.npz"), bin1=k["bin1"], bin2=k["bin2"],
                        count=k["count"], nbins=np.int64(k["nbins"]), weight=k["weight"])
    for fn in ("small_ice.npz", "small_kr.npz", "gm12878.npz", "kr_partial.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))
    print("gm12878 iterations", iters, "outliers", len(outl), "perchr outliers", len(outl_perchr))


if __name__ == "__main__":
    main()
