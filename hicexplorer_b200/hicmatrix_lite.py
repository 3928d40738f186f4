"""The slice of ``hicmatrix.HiCMatrix.hiCMatrix`` that the correction driver uses
(call sites hicexplorer/hicCorrectMatrix.py:603-616, 649, 660-698, 740-779), re-implemented on top of
``hdf5lite`` because hicmatrix / PyTables / cooler / h5py are third-party packages that are not part of the
reference tree and are not installable here.  Semantics follow SURVEY.md 8(b) and Appendix B (verified there
by reproducing the reference's golden ICE file bit-exactly).

Only host-side bookkeeping lives here (bin intervals, masks, file I/O); the matrix arithmetic is on the GPU.

File layouts (reference: docs/content/file-formats.rst:9-147):
  .h5    /matrix/{data,indices,indptr,shape} (upper triangle CSR), /intervals/{chr_list,start_list,end_list,
         extra_list}, /nan_bins, /correction_factors
  .cool  cooler schema: /bins/{chrom,start,end[,weight]}, /chroms/{name,length}, /indexes/{bin1_offset,
         chrom_offset}, /pixels/{bin1_id,bin2_id,count}
The writer emits uncompressed datasets (the reference writes Blosc / deflate compressed ones; readers do not care).
"""
import logging
from collections import OrderedDict

import numpy as np
from scipy import sparse

from . import hdf5lite

log = logging.getLogger(__name__)


def _to_str(x):
    return x.decode("utf-8") if isinstance(x, (bytes, np.bytes_)) else str(x)


class hiCMatrix(object):
    def __init__(self, pMatrixFile=None, pChrnameList=None):
        self._matrix = None          # full symmetric CSR (built lazily from _upper)
        self._upper = None           # upper triangle as stored in the file, while nothing has modified the matrix
        self._file_layout = None     # corrected upper triangle in the original frame, ready to be written
        self._upper_keep = None      # bins of _upper that are still part of the matrix (a chromosome subset in file
                                     # order was selected); None = all.  The selection itself is left to the device.
        self.value_dtype = np.float64  # dtype of /matrix/data (hicNormalize writes the float32 it computes in)
        self.cut_intervals = []
        self.nan_bins = np.array([], dtype=np.int64)
        self.correction_factors = None
        self.orig_bin_ids = []
        self.orig_cut_intervals = []
        self.interval_trees = OrderedDict()
        self.chrBinBoundaries = OrderedDict()
        self.matrixFileName = pMatrixFile
        if pMatrixFile is not None:
            fname = str(pMatrixFile).partition("::")[0]
            if fname.endswith(".cool") or fname.endswith(".mcool"):
                self._load_cool(pMatrixFile, pChrnameList)
            else:
                self._load_h5(pMatrixFile)
            self._index_intervals()

    @property
    def matrix(self):
        """full symmetric scipy CSR, like hicmatrix keeps it; symmetrised on first use"""
        if self._matrix is None and self._upper is not None:
            full = self._fill_lower(self._upper)
            if self._upper_keep is not None:
                keep = np.flatnonzero(self._upper_keep)
                full = sparse.csr_matrix(full[keep, :][:, keep])
                full.sort_indices()
                self._upper = None           # it still has the file's bins: no longer the same matrix
                self._upper_keep = None
            self._matrix = full
        return self._matrix

    @matrix.setter
    def matrix(self, value):
        self._matrix = value
        self._upper = None
        self._upper_keep = None
        self._file_layout = None

    def upper_for_device(self):
        """the file's upper triangle if the matrix is still exactly what was loaded (the device can then build the
        symmetric store itself and only half the bytes are uploaded), else None.  When a chromosome subset has been
        selected (``upper_drop_mask()`` is not None) the triangle still has the file's bins and the device deletes the
        others."""
        return self._upper

    def upper_drop_mask(self):
        """uint8 mask over the bins of ``upper_for_device()``: 1 = not part of the matrix any more; None = keep all"""
        if self._upper is None or self._upper_keep is None:
            return None
        return (~self._upper_keep).astype(np.uint8)

    def install_file_layout(self, upper, nan_bins, correction_factors):
        """take a corrected matrix that is already in file layout (upper triangle, original frame): what
        setMatrixValues + setCorrectionFactors + restoreMaskedBins would produce, computed on the device"""
        self._file_layout = upper
        self.nan_bins = np.asarray(nan_bins, dtype=np.int64)
        self.correction_factors = correction_factors

    # ---- loading ------------------------------------------------------------------------------------
    def _load_h5(self, path):
        f = hdf5lite.File(path)
        shape = tuple(int(v) for v in f["matrix/shape"].read())
        # the three big arrays stay views of the mapped file when it stores them plainly (files written by this package);
        # they are only ever read (uploaded, or copied by the symmetric fill)
        up = sparse.csr_matrix((f["matrix/data"].read(copy=False), f["matrix/indices"].read(copy=False),
                                f["matrix/indptr"].read(copy=False)), shape=shape)
        self._set_upper(up)
        chrom = f["intervals/chr_list"].read()
        start = f["intervals/start_list"].read()
        end = f["intervals/end_list"].read()
        extra = f["intervals/extra_list"].read()
        self.cut_intervals = [(_to_str(c), int(s), int(e), float(x)) for c, s, e, x in zip(chrom, start, end, extra)]
        self.nan_bins = f["nan_bins"].read() if "nan_bins" in f else np.array([], dtype=np.int64)
        self.correction_factors = f["correction_factors"].read() if "correction_factors" in f else None

    def _load_cool(self, path, chrnames):
        # cooler URIs: "file.mcool::/resolutions/10000" selects one matrix of a multi-resolution file
        fname, _sep, group = str(path).partition("::")
        f = hdf5lite.File(fname)
        if group.strip("/"):
            f = f[group]
        names = [_to_str(c) for c in f["chroms/name"].read()]
        chrom_id = f["bins/chrom"].read()
        start = f["bins/start"].read()
        end = f["bins/end"].read()
        n = len(start)
        b2 = f["pixels/bin2_id"].read(copy=False)
        cnt = f["pixels/count"].read(copy=False)
        up = None
        if "indexes" in f and "bin1_offset" in f["indexes"].keys():
            # the pixel table is sorted by (bin1, bin2) and indexes/bin1_offset is its row pointer (cooler schema): the
            # CSR exists already, no COO -> CSR conversion of a multi-gigabyte table
            offs = f["indexes/bin1_offset"].read()
            if len(offs) == n + 1 and offs[0] == 0 and offs[-1] == len(cnt) and np.all(np.diff(offs) >= 0):
                up = sparse.csr_matrix((cnt, b2, offs), shape=(n, n))
        if up is None:
            up = sparse.csr_matrix((np.asarray(cnt), (f["pixels/bin1_id"].read(), np.asarray(b2))), shape=(n, n))
        self._set_upper(up)
        self.cut_intervals = [(names[c], int(s), int(e), 1.0) for c, s, e in zip(chrom_id, start, end)]
        if "weight" in f["bins"].keys():
            self.correction_factors = f["bins/weight"].read()
        self.nan_bins = np.array([], dtype=np.int64)
        if chrnames:
            self._index_intervals()
            self.reorderChromosomes(chrnames)

    def _set_upper(self, up):
        up = sparse.csr_matrix(up)
        if hdf5lite.csr_is_canonical(up):    # scipy's one-thread test of the same property: 0.43 s on a 10 kb matrix
            up.has_canonical_format = True
        try:
            up.sum_duplicates()              # also sorts the columns of every row (no-op for canonical input)
        except ValueError:                   # read-only views of a mapped file that is not canonical: work on a copy
            up = up.copy()
            up.sum_duplicates()
        # entries below the diagonal?  With sorted rows only the first column of each row has to be looked at: O(n)
        # instead of a COO copy of a multi-gigabyte matrix
        rows = np.flatnonzero(np.diff(up.indptr))
        has_lower = bool(len(rows)) and bool(np.any(up.indices[up.indptr[rows]] < rows))
        if has_lower:                        # a file that stores both triangles: nothing to mirror
            self._matrix = up
            self._upper = None
        else:
            self._matrix = None
            self._upper = up

    @staticmethod
    def _fill_lower(up):
        """the file holds the upper triangle; in memory the matrix is full symmetric"""
        up = sparse.csr_matrix(up)
        full = sparse.csr_matrix(up + sparse.triu(up, 1).T)
        full.sort_indices()
        return full

    def _index_intervals(self):
        self.chrBinBoundaries = OrderedDict()
        self.interval_trees = OrderedDict()
        prev = None
        for i, iv in enumerate(self.cut_intervals):
            c = iv[0]
            if c != prev:
                if prev is not None:
                    self.chrBinBoundaries[prev] = (self.chrBinBoundaries[prev][0], i)
                if c in self.chrBinBoundaries:
                    raise ValueError("chromosome %s is not contiguous in the bin table" % c)
                self.chrBinBoundaries[c] = (i, i)
                self.interval_trees[c] = None
                prev = c
        if prev is not None:
            self.chrBinBoundaries[prev] = (self.chrBinBoundaries[prev][0], len(self.cut_intervals))

    # ---- queries --------------------------------------------------------------------------------------
    def getChrNames(self):
        return list(self.chrBinBoundaries)

    def getChrBinRange(self, chrname):
        return self.chrBinBoundaries[chrname]

    def chromosome_offsets(self):
        offs = [0]
        for c in self.chrBinBoundaries:
            offs.append(self.chrBinBoundaries[c][1])
        return np.asarray(offs, dtype=np.int64)

    # ---- restructuring ----------------------------------------------------------------------------------
    def reorderChromosomes(self, new_chr_order):
        new_chr_order = [_to_str(c) for c in new_chr_order]
        self.restoreMaskedBins()
        for c in new_chr_order:
            if c not in self.chrBinBoundaries:
                raise ValueError("Chromosome name '%s' not in matrix." % c)
        order = np.concatenate([np.arange(*self.chrBinBoundaries[c]) for c in new_chr_order]).astype(np.int64)
        if (self._upper is not None and self._matrix is None and self._upper_keep is None and len(order)
                and np.all(np.diff(order) > 0)):
            # a subset of the chromosomes in file order: the upper triangle stays an upper triangle, so the selection is
            # a bin deletion -- recorded here, carried out on the device (hb_csr_compact) right after the upload
            keep = np.zeros(self._upper.shape[0], dtype=bool)
            keep[order] = True
            self._upper_keep = keep
        else:
            self.matrix = sparse.csr_matrix(self.matrix[order, :][:, order])
            self.matrix.sort_indices()
        self.cut_intervals = [self.cut_intervals[i] for i in order]
        if self.correction_factors is not None:
            self.correction_factors = np.asarray(self.correction_factors).ravel()[order]
        pos = {old: new for new, old in enumerate(order)}
        self.nan_bins = np.array(sorted(pos[b] for b in self.nan_bins if b in pos), dtype=np.int64)
        self._index_intervals()

    def keepOnlyTheseChr(self, chromosome_list):
        """hicmatrix keepOnlyTheseChr: drop every chromosome that is not listed; unlike reorderChromosomes the FILE order
        of the chromosomes is kept (what hicTransform --chromosomes uses, hicTransform.py:156)"""
        wanted = set(_to_str(c) for c in chromosome_list)
        for c in wanted:
            if c not in self.chrBinBoundaries:
                raise ValueError("Chromosome name '%s' not in matrix." % c)
        self.reorderChromosomes([c for c in self.chrBinBoundaries if c in wanted])
        return self.matrix

    def maskBins(self, bin_ids=None):
        """delete rows+columns; a second call first restores, then removes the union"""
        if bin_ids is None or len(bin_ids) == 0:
            return
        bin_ids = np.asarray(bin_ids, dtype=np.int64)
        if len(self.orig_bin_ids) > 0:
            # ids refer to the current (already masked) frame: translate, restore, remove the union
            absolute = np.asarray(self.orig_bin_ids)[bin_ids]
            previous = self._masked_ids
            self.restoreMaskedBins()
            bin_ids = np.union1d(absolute, previous)
        n = self.matrix.shape[0]
        keep = np.setdiff1d(np.arange(n), bin_ids)
        self._masked_ids = np.asarray(sorted(set(bin_ids.tolist())), dtype=np.int64)
        self.orig_bin_ids = keep
        self.orig_cut_intervals = list(self.cut_intervals)
        self._orig_n = n
        self.matrix = sparse.csr_matrix(self.matrix[keep, :][:, keep])
        self.cut_intervals = [self.cut_intervals[i] for i in keep]
        if self.correction_factors is not None:
            self.correction_factors = np.asarray(self.correction_factors).ravel()[keep]
        self.nan_bins = np.array([], dtype=np.int64)
        self._index_intervals()

    _masked_ids = np.array([], dtype=np.int64)

    def apply_mask_bookkeeping(self, kept_ids, matrix=None):
        """Record that only ``kept_ids`` (ids in the CURRENT frame) survive -- used when the deletion itself
        happened on the GPU; optionally install the already-compacted matrix."""
        kept_ids = np.asarray(kept_ids, dtype=np.int64)
        if len(self.orig_bin_ids) > 0:
            raise RuntimeError("nested device masks are not supported; restore first")
        n = len(self.cut_intervals)
        self._masked_ids = np.setdiff1d(np.arange(n), kept_ids)
        self.orig_bin_ids = kept_ids
        self.orig_cut_intervals = list(self.cut_intervals)
        self._orig_n = n
        self.cut_intervals = [self.cut_intervals[i] for i in kept_ids]
        if matrix is not None:
            self.matrix = matrix
        else:
            self.matrix = sparse.csr_matrix(self.matrix[kept_ids, :][:, kept_ids])
        if self.correction_factors is not None:
            self.correction_factors = np.asarray(self.correction_factors).ravel()[kept_ids]
        self._index_intervals()

    def restoreMaskedBins(self):
        """re-insert the masked bins as empty rows/columns; they become nan_bins"""
        if len(self.orig_bin_ids) == 0:
            return
        n = self._orig_n
        kept = np.asarray(self.orig_bin_ids, dtype=np.int64)
        coo = self.matrix.tocoo()
        self.matrix = sparse.csr_matrix((coo.data, (kept[coo.row], kept[coo.col])), shape=(n, n))
        self.matrix.sort_indices()
        self.cut_intervals = list(self.orig_cut_intervals)
        self.nan_bins = np.setdiff1d(np.arange(n), kept).astype(np.int64)
        if self.correction_factors is not None:
            cf = np.asarray(self.correction_factors, dtype=np.float64)
            shape = (n,) if cf.ndim == 1 else (n, 1)
            full = np.full(shape, np.nan)
            full[kept] = cf.reshape((len(kept),) + shape[1:])
            self.correction_factors = full
        self.orig_bin_ids = []
        self.orig_cut_intervals = []
        self._masked_ids = np.array([], dtype=np.int64)
        self._index_intervals()

    def diagflat(self, value=np.nan):
        m = self.matrix.tolil()
        m.setdiag(value)
        self.matrix = m.tocsr()
        if value == 0:
            self.matrix.eliminate_zeros()
        return self.matrix

    def printchrtoremove(self, to_remove, label="Number of poor regions to remove", restore_masked_bins=True):
        cnt = OrderedDict()
        for idx in to_remove:
            c = self.cut_intervals[idx][0]
            cnt[c] = cnt.get(c, 0) + 1
        log.info('{}: {} {}'.format(label, len(to_remove), dict(cnt)))

    def setMatrixValues(self, newMatrix):
        assert self.matrix.shape == newMatrix.shape, "Given matrix has different shape. New values need to have the same shape as previous matrix."
        self.matrix = sparse.csr_matrix(newMatrix)

    def setCorrectionFactors(self, correction_factors):
        n_bins = self._matrix.shape[0] if self._matrix is not None else len(self.cut_intervals)   # no host symmetrisation
        assert len(correction_factors) == n_bins, "length of correction factors and length of matrix are different."
        self.correction_factors = correction_factors

    # ---- saving ----------------------------------------------------------------------------------------------
    def save(self, pMatrixName, pSymmetric=True, pApplyCorrection=False):
        if self._file_layout is None:
            self.restoreMaskedBins()
        if str(pMatrixName).endswith(".cool"):
            self._save_cool(pMatrixName)
        else:
            if not str(pMatrixName).endswith(".h5"):
                pMatrixName = str(pMatrixName) + ".h5"
            self._save_h5(pMatrixName)

    def _upper_of_matrix(self):
        if self._file_layout is not None:
            return self._file_layout
        up = sparse.triu(self.matrix, k=0, format="csr")
        up.eliminate_zeros()
        up.sort_indices()
        return up

    def _save_h5(self, path):
        up = self._upper_of_matrix()
        # chunked + Blosc like the reference's files (PyTables CArray); HDF5LITE_COMPRESS=0 writes plain contiguous Arrays
        carray = ({"CLASS": "CARRAY", "VERSION": "1.1", "TITLE": ""} if hdf5lite._compression_enabled() else
                  {"CLASS": "ARRAY", "VERSION": "2.4", "TITLE": "", "FLAVOR": "numpy"})
        group = {"CLASS": "GROUP", "VERSION": "1.0", "TITLE": ""}
        w = hdf5lite.Writer({"TITLE": "HiCExplorer matrix", "CLASS": "GROUP", "VERSION": "1.0",
                             "PYTABLES_FORMAT_VERSION": "2.1"})
        w.group("/matrix", group)
        w.group("/intervals", group)
        # (no copies when the arrays already have the file's types: these are the multi-gigabyte payloads)
        w.dataset("/matrix/data", up.data.astype(self.value_dtype, copy=False), carray, filters="blosc")
        w.dataset("/matrix/indices", up.indices.astype(np.int32, copy=False), carray, filters="blosc")
        w.dataset("/matrix/indptr", up.indptr.astype(np.int32 if up.nnz < 2 ** 31 else np.int64, copy=False), carray, filters="blosc")
        w.dataset("/matrix/shape", np.asarray(up.shape, dtype=np.int64), carray, filters="blosc")
        chrom, start, end, extra = zip(*self.cut_intervals) if self.cut_intervals else ((), (), (), ())
        width = max([len(c) for c in chrom] + [1])
        w.dataset("/intervals/chr_list", np.asarray([c.encode() for c in chrom], dtype="S%d" % width), carray, filters="blosc")
        w.dataset("/intervals/start_list", np.asarray(start, dtype=np.int64), carray, filters="blosc")
        w.dataset("/intervals/end_list", np.asarray(end, dtype=np.int64), carray, filters="blosc")
        w.dataset("/intervals/extra_list", np.asarray(extra, dtype=np.float64), carray, filters="blosc")
        if len(self.nan_bins):
            w.dataset("/nan_bins", np.asarray(self.nan_bins, dtype=np.int64), carray, filters="blosc")
        if self.correction_factors is not None:
            w.dataset("/correction_factors", np.asarray(self.correction_factors, dtype=np.float64), carray, filters="blosc")
        w.save(path)

    def _save_cool(self, path):
        up = self._upper_of_matrix().tocoo()
        chrom, start, end, _extra = zip(*self.cut_intervals)
        names = list(OrderedDict.fromkeys(chrom))
        ids = {c: i for i, c in enumerate(names)}
        chrom_id = np.asarray([ids[c] for c in chrom], dtype=np.int32)
        n = len(chrom)
        lengths = [max(e for c2, e in zip(chrom, end) if c2 == c) for c in names]
        bin1_offset = np.concatenate([[0], np.cumsum(np.bincount(up.row, minlength=n))]).astype(np.int64)
        chrom_offset = np.concatenate([[0], np.cumsum(np.bincount(chrom_id, minlength=len(names)))]).astype(np.int64)
        sizes = np.asarray(end, dtype=np.int64) - np.asarray(start, dtype=np.int64)
        w = hdf5lite.Writer({"format": "HDF5::Cooler", "format-version": np.int64(3), "bin-type": "fixed",
                             "bin-size": np.int64(np.median(sizes)) if n else np.int64(0), "storage-mode": "symmetric-upper",
                             "nbins": np.int64(n), "nchroms": np.int64(len(names)), "nnz": np.int64(up.nnz),
                             "sum": np.float64(up.data.sum()), "generated-by": "hicexplorer_b200"})
        width = max([len(c) for c in names] + [1])
        w.dataset("/chroms/name", np.asarray([c.encode() for c in names], dtype="S%d" % width), filters="gzip-shuffle")
        w.dataset("/chroms/length", np.asarray(lengths, dtype=np.int32), filters="gzip-shuffle")
        w.dataset("/bins/chrom", chrom_id, enum=OrderedDict((c, i) for i, c in enumerate(names)), filters="gzip-shuffle")
        w.dataset("/bins/start", np.asarray(start, dtype=np.int32), filters="gzip-shuffle")
        w.dataset("/bins/end", np.asarray(end, dtype=np.int32), filters="gzip-shuffle")
        if self.correction_factors is not None:
            w.dataset("/bins/weight", np.asarray(self.correction_factors, dtype=np.float64).ravel(), filters="gzip-shuffle")
        w.dataset("/indexes/bin1_offset", bin1_offset, filters="gzip-shuffle")
        w.dataset("/indexes/chrom_offset", chrom_offset, filters="gzip-shuffle")
        w.dataset("/pixels/bin1_id", up.row.astype(np.int32), filters="gzip-shuffle")
        w.dataset("/pixels/bin2_id", up.col.astype(np.int32), filters="gzip-shuffle")
        w.dataset("/pixels/count", up.data.astype(self.value_dtype, copy=False), filters="gzip-shuffle")
        w.save(path)
