"""Host-side owner of a device-resident CSR block (``hb_csr`` in
include/hicbalance.h).  Everything numeric happens in libhicbalance.so; this
class only moves buffers and translates status codes.

The block is what the reference keeps as ``ma.matrix`` (a full symmetric scipy
CSR, hicCorrectMatrix.py:603-626), living in HBM as int64 ``indptr`` / int32
``indices`` / fp64 ``data``.
"""
import ctypes

import numpy as np
from scipy import sparse

from . import _lib
from ._lib import check, lib, ptr

KR_TOL = 1e-6
KR_DELTA = 0.1
KR_DELTA_UPPER = 3.0
KR_DIAG_EPS = 1e-5


# value types the library widens to fp64 on the device (hb_csr_create_typed): no fp64 copy of the matrix on the host
_VALUE_KINDS = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}


def canonical_csr(matrix, dtype=np.float64):
    """CSR with sorted, duplicate-free rows (a copy only when needed).  ``dtype=None`` keeps a value type the device
    can widen itself (float64, float32, int32, int64) and converts anything else to float64."""
    m = matrix if sparse.isspmatrix_csr(matrix) else sparse.csr_matrix(matrix)
    want = dtype if dtype is not None else (m.dtype if m.dtype in _VALUE_KINDS else np.float64)
    if m.dtype != want:
        m = m.astype(want)
    if not m.has_canonical_format:
        m = m.copy() if m is matrix else m
        m.sum_duplicates()
    return m


def parallel_copy(arr, threads=8):
    """np.copy of a large 1-D array on a few threads (numpy releases the GIL for the slices): the first touch of the
    fresh pages and the copy itself both scale, a multi-gigabyte index array comes out in a fraction of the time."""
    arr = np.ascontiguousarray(arr)
    if arr.nbytes < (64 << 20) or arr.ndim != 1:
        return arr.copy()
    from concurrent.futures import ThreadPoolExecutor
    out = np.empty_like(arr)
    step = -(-len(arr) // threads)

    def part(k):
        out[k * step:(k + 1) * step] = arr[k * step:(k + 1) * step]

    with ThreadPoolExecutor(max_workers=threads) as pool:
        list(pool.map(part, range(threads)))
    return out


def percentile_position(count, q):
    """Where numpy's default ('linear') percentile looks in the sorted sample: index of the lower neighbour and the
    interpolation weight, computed with numpy's own expressions (virtual index = (count - 1) * (q / 100))."""
    quantile = np.true_divide(np.float64(q), 100)
    virtual = (count - 1) * quantile
    previous = np.floor(virtual)
    gamma = virtual - previous
    k = int(previous)
    if k >= count - 1:
        return count - 1, np.float64(0.0)
    if k < 0:
        return 0, np.float64(0.0)
    return k, gamma


def lerp_like_numpy(a, b, t):
    """numpy's _lerp: a + (b - a) * t, switched to b - (b - a) * (1 - t) for t >= 0.5"""
    a, b, t = np.float64(a), np.float64(b), np.float64(t)
    diff = b - a
    return b - diff * (1 - t) if t >= 0.5 else a + diff * t


class DeviceCSR(object):
    def __init__(self, handle, device):
        self._h = ctypes.c_void_p(handle) if not isinstance(handle, ctypes.c_void_p) else handle
        self.device = device

    # ---- construction -------------------------------------------------------
    @classmethod
    def from_scipy(cls, matrix, device=0, row_range=None):
        """Upload ``matrix`` (square, symmetric, canonical CSR).  ``row_range`` =
        (row0, row1) uploads only that block of rows (sharded runs)."""
        return cls.from_scipy_as_uploaded(matrix, device=device, row_range=row_range)[0]

    @classmethod
    def from_scipy_as_uploaded(cls, matrix, device=0, row_range=None):
        """``from_scipy`` that also returns the host matrix whose pattern the store has.  A CSR input is uploaded as it
        is -- the library checks sorted, duplicate-free rows on the device in a few milliseconds
        (hb_csr_create validation), scipy's host-side check of a genome-wide matrix takes the better part of a second --
        and only a rejected block is canonicalised on the host (a copy) and uploaded again."""
        _lib.require_device()
        m = matrix if sparse.isspmatrix_csr(matrix) else sparse.csr_matrix(matrix)
        if m.dtype not in _VALUE_KINDS:
            m = m.astype(np.float64)
        if m.shape[0] != m.shape[1]:
            raise ValueError("matrix must be square")
        n = m.shape[0]
        row0, row1 = (0, n) if row_range is None else row_range
        for attempt in (0, 1):
            indptr = np.ascontiguousarray(m.indptr[row0:row1 + 1], dtype=np.int64)
            nnz = int(indptr[-1] - indptr[0])
            idx_i64 = 1 if m.indices.dtype == np.int64 else 0
            indices = np.ascontiguousarray(m.indices, dtype=np.int64 if idx_i64 else np.int32)
            data = np.ascontiguousarray(m.data)
            h = ctypes.c_void_p()
            rc = lib().hb_csr_create_typed(ctypes.byref(h), row1 - row0, n, row0, nnz, ptr(indptr), ptr(indices), idx_i64,
                                           ptr(data), _VALUE_KINDS[data.dtype], device)
            if rc == _lib.HB_ERR_INVALID_ARG and attempt == 0 and "ascending" in _lib.last_error():
                m = m.copy() if m is matrix else m
                m.sum_duplicates()               # sorts the columns and merges duplicates
                continue
            check(rc)
            return cls(h, device), m

    @classmethod
    def from_upper(cls, upper, device=0):
        """Upload only the upper triangle (file layout) and build the full symmetric store on the device."""
        _lib.require_device()
        m = canonical_csr(upper, dtype=None)
        n = m.shape[0]
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        data = np.ascontiguousarray(m.data)
        h = ctypes.c_void_p()
        check(lib().hb_csr_create_upper_typed(ctypes.byref(h), n, int(m.nnz), ptr(indptr), ptr(indices), ptr(data),
                                              _VALUE_KINDS[data.dtype], device))
        return cls(h, device)

    @classmethod
    def wrap_device(cls, n_rows, n_cols, row0, nnz, d_indptr, d_indices, d_data, device=0):
        """Zero-copy view of arrays already in HBM (raw device pointers, e.g. ``tensor.data_ptr()``)."""
        _lib.require_device()
        h = ctypes.c_void_p()
        check(lib().hb_dev_csr_wrap(ctypes.byref(h), n_rows, n_cols, row0, nnz, d_indptr, d_indices, d_data, device))
        return cls(h, device)

    @classmethod
    def synthetic(cls, chrom_bins, seed=0, band=2000, d0=200.0, gamma=1.0, lambda0=40.0, n_trans=8,
                  low_frac=0.02, empty_frac=0.005, lambda_trans=2.0, row_range=None, device=0):
        """Deterministic synthetic Hi-C matrix generated directly in HBM."""
        _lib.require_device()
        cb = np.ascontiguousarray(chrom_bins, dtype=np.int64)
        n = int(cb.sum())
        row0, row1 = (0, n) if row_range is None else row_range
        p = _lib.SynthParams(seed=seed, nchrom=len(cb), band=band, n_trans=n_trans, reserved=0, d0=d0, gamma=gamma,
                             lambda0=lambda0, low_frac=low_frac, empty_frac=empty_frac, lambda_trans=lambda_trans)
        h = ctypes.c_void_p()
        check(lib().hb_synth_create(ctypes.byref(h), ctypes.byref(p), ptr(cb), row0, row1 - row0, device))
        return cls(h, device)

    def close(self):
        if self._h is not None and self._h.value:
            lib().hb_csr_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- introspection ------------------------------------------------------
    @property
    def handle(self):
        return self._h

    def dims(self):
        a, b, c, d = (ctypes.c_int64() for _ in range(4))
        check(lib().hb_csr_shape(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return a.value, b.value, c.value, d.value

    @property
    def n_rows(self):
        return self.dims()[0]

    @property
    def n(self):
        return self.dims()[1]

    @property
    def row0(self):
        return self.dims()[2]

    @property
    def nnz(self):
        return self.dims()[3]

    def device_pointers(self):
        a, b, c = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        check(lib().hb_dev_csr_pointers(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def sync(self):
        check(lib().hb_sync(self._h))

    def download(self, values_only=False):
        """Host copy of the block as a scipy CSR (rows = block rows, columns = all bins)."""
        n_rows, n_cols, _row0, nnz = self.dims()
        data = np.empty(nnz, dtype=np.float64)
        if values_only:
            check(lib().hb_csr_download(self._h, None, None, ptr(data)))
            return data
        indptr = np.empty(n_rows + 1, dtype=np.int64)
        indices = np.empty(nnz, dtype=np.int32)
        check(lib().hb_csr_download(self._h, ptr(indptr), ptr(indices), ptr(data)))
        return sparse.csr_matrix((data, indices, indptr), shape=(n_rows, n_cols))

    # ---- pre-filter ---------------------------------------------------------
    def set_segments(self, offsets):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        check(lib().hb_set_segments(self._h, len(off) - 1, ptr(off)))
        self._segments = len(off) - 1

    def keep_cis(self):
        check(lib().hb_csr_keep_cis(self._h))

    def scrub(self):
        """NaN -> 0, Inf -> 0 on the device; returns (#nan, #inf)."""
        counts = np.zeros(2, dtype=np.int64)
        check(lib().hb_scrub(self._h, ptr(counts)))
        return int(counts[0]), int(counts[1])

    def pack_values(self):
        """Lossless narrow copy of the values for the streaming kernels; returns 0 (none), 1 (uint16) or 2 (float)."""
        kind = ctypes.c_int()
        check(lib().hb_csr_pack_values(self._h, ctypes.byref(kind)))
        return kind.value

    def row_stats(self, cis_only=False):
        """(row sums, diagonal) of the block's rows; ``cis_only``: sums restricted to each row's own segment"""
        n_rows = self.n_rows
        rs = np.empty(n_rows, dtype=np.float64)
        dg = np.empty(n_rows, dtype=np.float64)
        fn = lib().hb_row_stats_cis if cis_only else lib().hb_row_stats
        check(fn(self._h, ptr(rs), ptr(dg)))
        return rs, dg

    def mad_filter(self, lo, hi, want_z=False):
        """(mask uint8[n], zscore or None, median[nseg], mad[nseg])."""
        n = self.n
        mask = np.zeros(n, dtype=np.uint8)
        z = np.empty(n, dtype=np.float64) if want_z else None
        nseg = max(1, self._nseg())
        med = np.empty(nseg, dtype=np.float64)
        mad = np.empty(nseg, dtype=np.float64)
        check(lib().hb_mad_filter(self._h, float(lo), float(hi), ptr(mask), ptr(z), ptr(med), ptr(mad)))
        return mask, z, med, mad

    _segments = 1

    def _nseg(self):
        return self._segments

    def compact(self, remove_mask):
        m = np.ascontiguousarray(remove_mask, dtype=np.uint8)
        if len(m) != self.n:
            raise ValueError("mask length must equal the bin count")
        check(lib().hb_csr_compact(self._h, ptr(m)))

    def merge_bins(self, bin_map, n_new):
        """Sum groups of consecutive bins (rows and columns): bin_map[i] = merged id of bin i or -1 (dropped)."""
        m = np.ascontiguousarray(bin_map, dtype=np.int32)
        if len(m) != self.n:
            raise ValueError("bin map length must equal the bin count")
        kept = m[m >= 0]
        sizes = np.bincount(kept, minlength=int(n_new)) if len(kept) else np.zeros(0, dtype=np.int64)
        if len(sizes) and sizes.max() > 32:
            # The kernel merges at most 32 bins per group (one lane per old row).  Larger groups are merged in two exact
            # steps: first sub-groups of <= 32 consecutive bins, then the sub-groups of every target group (sums of
            # sums; the upper-triangle rule of the diagonal cells composes, see csrc/hb_merge.cu).
            if sizes.max() > 1024:
                raise ValueError("at most 1024 bins per merged bin")
            first = np.full(len(m), -1, dtype=np.int32)
            second = []
            pos = np.flatnonzero(m >= 0)
            starts = np.flatnonzero(np.r_[True, m[pos][1:] != m[pos][:-1]])
            ends = np.r_[starts[1:], len(pos)]
            nxt = 0
            for a, b in zip(starts, ends):
                ids = pos[a:b]
                if len(ids) and (ids[-1] - ids[0] + 1 != len(ids)):
                    raise ValueError("the bins of a group must be consecutive")
                sub = np.arange(len(ids)) // 32
                first[ids] = nxt + sub
                n_sub = int(sub[-1]) + 1
                second.extend([int(m[ids[0]])] * n_sub)
                nxt += n_sub
            check(lib().hb_csr_merge_bins(self._h, ptr(first), int(nxt)))
            m = np.ascontiguousarray(second, dtype=np.int32)
        check(lib().hb_csr_merge_bins(self._h, ptr(m), int(n_new)))
        self._segments = 1

    def keep_rows(self, row_begin, row_end):
        """full matrix -> row block [row_begin, row_end) (sharded runs that uploaded the whole upper triangle)"""
        check(lib().hb_csr_keep_rows(self._h, int(row_begin), int(row_end)))

    def download_indptr(self):
        """row pointer of the block (int64[n_rows + 1]); the entries stay on the device"""
        ip = np.empty(self.n_rows + 1, dtype=np.int64)
        check(lib().hb_csr_download(self._h, ptr(ip), None, None))
        return ip

    def symmetrize_upper(self):
        """upper-triangular store -> full symmetric store (M + triu(M, 1)^T), on the device"""
        check(lib().hb_csr_symmetrize_upper(self._h))

    def shifted_copy(self, dr, dc):
        """new store C with C[i + dr][j + dc] = self[i][j] (targets outside the frame dropped)"""
        h = ctypes.c_void_p()
        check(lib().hb_csr_shifted_copy(ctypes.byref(h), self._h, int(dr), int(dc)))
        return DeviceCSR(h, self.device)

    def keep_upper(self):
        check(lib().hb_csr_keep_upper(self._h))

    def diag_zero(self):
        check(lib().hb_diag_zero(self._h))

    def symmetry_ratio(self):
        r = ctypes.c_double()
        rc = lib().hb_symmetry(self._h, ctypes.byref(r))
        if rc == _lib.HB_ERR_NOT_SYMMETRIC:      # row blocks: any asymmetry is reported as ratio 1
            return r.value
        check(rc)
        return r.value

    def symmetry_ratio_sharded(self, dist, rank, world):
        """iterativeCorrection.py:35 on row blocks.  The exact-symmetry screen first (one tiny allreduce); only a matrix
        that is not identical to its transpose pays for the exact ratio: every rank sends each entry to the owner of its
        column (one all-to-all of (row, column, value) triplets over NCCL), the owners sum |a_ij - a_ji|."""
        import torch
        r = self.symmetry_ratio()
        if r == 0.0:
            return 0.0
        n_rows, n_cols, row0, _nnz = self.dims()
        span = torch.zeros(world + 1, dtype=torch.int64, device="cuda")
        span[rank] = row0
        if rank == world - 1:
            span[world] = row0 + n_rows
        dist.all_reduce(span)
        bounds = span.cpu().numpy()
        L = lib()
        counts = np.zeros(world, dtype=np.int64)
        cnt = ctypes.c_int64()
        for q in range(world):
            check(L.hb_dev_extract_cols(self._h, int(bounds[q]), int(bounds[q + 1]), ctypes.byref(cnt), None, None, None))
            counts[q] = cnt.value
        total = int(counts.sum())
        rows = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
        cols = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
        vals = torch.empty(max(total, 1), dtype=torch.float64, device="cuda")
        off = np.concatenate([[0], np.cumsum(counts)])
        for q in range(world):
            if counts[q]:
                a = int(off[q])
                check(L.hb_dev_extract_cols(self._h, int(bounds[q]), int(bounds[q + 1]), ctypes.byref(cnt),
                                            ctypes.c_void_p(rows.data_ptr() + 4 * a), ctypes.c_void_p(cols.data_ptr() + 4 * a),
                                            ctypes.c_void_p(vals.data_ptr() + 8 * a)))
        send = torch.from_numpy(counts).cuda()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        recv_counts = recv.cpu().numpy()
        n_in = int(recv_counts.sum())
        got = []
        for t in (rows, cols, vals):
            buf = torch.empty(max(n_in, 1), dtype=t.dtype, device="cuda")
            dist.all_to_all_single(buf[:n_in], t[:total], [int(c) for c in recv_counts], [int(c) for c in counts])
            got.append(buf)
        torch.cuda.current_stream().synchronize()   # the collectives ran on torch's stream, the look-up runs on the handle's
        part = ctypes.c_double()
        check(L.hb_dev_partner_sums(self._h, n_in, ctypes.c_void_p(got[0].data_ptr()), ctypes.c_void_p(got[1].data_ptr()),
                                    ctypes.c_void_p(got[2].data_ptr()), ctypes.byref(part)))
        sums = torch.tensor([part.value, float(self.value_stats()[0])], dtype=torch.float64, device="cuda")
        dist.all_reduce(sums)
        num, den = float(sums[0].item()), abs(float(sums[1].item()))
        cells = float(n_cols) * float(n_cols)
        return (num / cells) / (den / cells) if den else float("inf")

    # ---- ICE ------------------------------------------------------------------
    def ice(self, max_iter=50, tolerance=1e-5):
        """Returns (bias[n], iterations[nseg], deviation[nseg]).  Raises OverflowIter."""
        n = self.n
        nseg = self._nseg()
        bias = np.empty(n, dtype=np.float64)
        iters = np.zeros(nseg, dtype=np.int32)
        dev = np.zeros(nseg, dtype=np.float64)
        check(lib().hb_ice_run(self._h, int(max_iter), float(tolerance), ptr(bias), ptr(iters), ptr(dev)))
        return bias, iters, dev

    def ice_scaled_values(self):
        """A_ij / (b_i b_j) in the store's entry order.  Raises OverflowIter / OverflowFinal."""
        out = np.empty(self.nnz, dtype=np.float64)
        mx = ctypes.c_double()
        check(lib().hb_ice_scaled_values(self._h, ptr(out), ctypes.byref(mx)))
        return out

    def ice_scaled_upper(self, orig_ids, n_orig):
        """Corrected matrix in file layout: upper triangle, zeros eliminated, in the original n_orig-bin frame."""
        orig = np.ascontiguousarray(orig_ids, dtype=np.int64)
        nnz = ctypes.c_int64()
        check(lib().hb_ice_scaled_upper(self._h, ptr(orig), int(n_orig), ctypes.byref(nnz), None, None, None))
        indptr = np.empty(n_orig + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.float64)
        check(lib().hb_ice_scaled_upper(self._h, ptr(orig), int(n_orig), ctypes.byref(nnz), ptr(indptr), ptr(indices),
                                        ptr(data)))
        return sparse.csr_matrix((data, indices, indptr), shape=(n_orig, n_orig))

    # ---- hicNormalize / hicSumMatrices (element-wise steps around the correction) --
    def value_stats(self, scrub=False):
        """(sum of the stored values in fp64, float32 min, float32 max) -- hicNormalize.py:70, 82-83"""
        s, mn, mx = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        check(lib().hb_csr_value_stats(self._h, int(bool(scrub)), ctypes.byref(s), ctypes.byref(mn), ctypes.byref(mx)))
        return s.value, mn.value, mx.value

    def normalize(self, mode, a=0.0, b=1.0, threshold=0.0, scrub_first=True):
        """float32 normalisation in place (mode: 'norm_range' (f-a)/b, 'divide' f/a, 'multiply' f*a, 'cast')"""
        code = {"norm_range": 0, "divide": 1, "multiply": 2, "cast": 3}[mode]
        check(lib().hb_csr_normalize(self._h, code, float(a), float(b), float(threshold), int(bool(scrub_first))))

    def eliminate_zeros(self):
        check(lib().hb_csr_eliminate_zeros(self._h))

    def add(self, other):
        """self <- self + other (same shape); exact zeros dropped like scipy's csr + csr"""
        check(lib().hb_csr_add(self._h, other.handle))

    # ---- hicTransform obs/exp -------------------------------------------------------
    def distance_stats(self):
        """(sum, count) of the stored non-zero values per genomic distance (per segment: slot seg_off[s] + d)"""
        n = self.n
        sums = np.zeros(n, dtype=np.float64)
        counts = np.zeros(n, dtype=np.int64)
        check(lib().hb_distance_stats(self._h, ptr(sums), ptr(counts)))
        return sums, counts

    def obs_exp_apply(self, expected, index_mode, out_mode, row_sum=None, seg_total=None):
        ex = np.ascontiguousarray(expected, dtype=np.float64)
        rs = None if row_sum is None else np.ascontiguousarray(row_sum, dtype=np.float64)
        tot = None if seg_total is None else np.ascontiguousarray(seg_total, dtype=np.float64)
        check(lib().hb_obs_exp_apply(self._h, ptr(ex), int(index_mode), int(out_mode), ptr(rs), ptr(tot)))

    # ---- hicTransform covariance / pearson -----------------------------------------
    def block_correlation(self, lo, hi, pearson=False):
        """np.cov / np.corrcoef of the dense diagonal block [lo, hi) (hicTransform.py:105-113, 236) as a scipy CSR of
        its non-zero cells: shape (hi - lo, n), columns in the frame of the whole matrix"""
        import scipy.sparse as sp
        lo, hi = int(lo), int(hi)
        nnz = ctypes.c_int64()
        check(lib().hb_block_correlation(self._h, lo, hi, int(bool(pearson)), ctypes.byref(nnz), None, None, None))
        indptr = np.zeros(hi - lo + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.float64)
        if hi > lo:
            check(lib().hb_block_correlation(self._h, lo, hi, int(bool(pearson)), ctypes.byref(nnz), ptr(indptr), ptr(indices),
                                             ptr(data)))
        return sp.csr_matrix((data, indices, indptr), shape=(hi - lo, self.n))

    # ---- optional steps around ICE (--transCutoff / --inflationCutoff) ------------
    def trans_percentile(self, chr_offsets, q):
        """np.percentile(values stored between chromosomes, q) -- what hicmatrix truncTrans computes
        (hicCorrectMatrix.py:695-698) -- from two exact order statistics selected on the device; None when the
        matrix has no trans entries."""
        offs = np.ascontiguousarray(chr_offsets, dtype=np.int64)
        nchr = len(offs) - 1
        count = ctypes.c_int64()
        check(lib().hb_trans_order_stats(self._h, ptr(offs), nchr, 0, ctypes.byref(count), None, None))
        if count.value == 0:
            return None
        k, gamma = percentile_position(count.value, q)
        a, b = ctypes.c_double(), ctypes.c_double()
        check(lib().hb_trans_order_stats(self._h, ptr(offs), nchr, k, ctypes.byref(count), ctypes.byref(a), ctypes.byref(b)))
        return lerp_like_numpy(a.value, b.value, gamma)

    def trans_clip(self, chr_offsets, max_inter):
        offs = np.ascontiguousarray(chr_offsets, dtype=np.int64)
        clipped = ctypes.c_int64()
        check(lib().hb_trans_clip(self._h, ptr(offs), len(offs) - 1, float(max_inter), ctypes.byref(clipped)))
        return clipped.value

    def ice_scaled_row_sums(self):
        out = np.empty(self.n_rows, dtype=np.float64)
        check(lib().hb_ice_scaled_row_sums(self._h, ptr(out)))
        return out

    # ---- Knight-Ruiz ------------------------------------------------------------
    def kr(self, tol=KR_TOL, delta=KR_DELTA, Delta=KR_DELTA_UPPER, diag_eps=KR_DIAG_EPS, max_outer=100000):
        """Returns (x[n], outer iterations, mat-vecs, residual norm)."""
        x = np.empty(self.n, dtype=np.float64)
        outer, mv, res = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
        check(lib().hb_kr_run(self._h, tol, delta, Delta, diag_eps, int(max_outer), ptr(x), ctypes.byref(outer),
                              ctypes.byref(mv), ctypes.byref(res)))
        return x, outer.value, mv.value, res.value

    def kr_rescale(self, x, diag_eps=KR_DIAG_EPS):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        f = ctypes.c_double()
        check(lib().hb_kr_rescale(self._h, diag_eps, ptr(x), ctypes.byref(f)))
        return x, f.value

    def kr_scaled_upper(self, x, diag_eps=KR_DIAG_EPS):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n_rows, n, _row0, _nnz = self.dims()
        nnz = ctypes.c_int64()
        check(lib().hb_kr_scaled_upper(self._h, diag_eps, ptr(x), ctypes.byref(nnz), None, None, None))
        indptr = np.empty(n_rows + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.float64)
        check(lib().hb_kr_scaled_upper(self._h, diag_eps, ptr(x), ctypes.byref(nnz), ptr(indptr), ptr(indices),
                                       ptr(data)))
        return sparse.csr_matrix((data, indices, indptr), shape=(n_rows, n))
