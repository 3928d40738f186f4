"""Device-resident correction pipelines: what ``hicCorrectMatrix correct`` does
between loading the matrix and saving it (hicexplorer/hicCorrectMatrix.py:612-757),
with the matrix uploaded ONCE and every stage (zero-bin mask, NaN/Inf scrub,
MAD filter, bin deletion, ICE / KR) running on the GPU.

Host code here only sequences C-ABI calls and does O(n) index bookkeeping.
"""
import logging

import numpy as np
from scipy import sparse

from . import _lib
from .store import KR_DIAG_EPS, DeviceCSR, canonical_csr

log = logging.getLogger(__name__)


def _segment_offsets_after(mask_removed, offsets):
    """chromosome offsets after deleting the bins flagged in ``mask_removed``."""
    kept_before = np.concatenate([[0], np.cumsum(~mask_removed.astype(bool))])
    return kept_before[np.asarray(offsets, dtype=np.int64)]


def ice_pipeline(matrix, filter_threshold, max_iter=500, tolerance=1e-5, chr_offsets=None, perchr=False,
                 skip_diagonal=False, device=0, want_matrix=True, extra_mask_fn=None, upper_input=False,
                 file_layout=False, trans_cutoff=None, clip_trans=False, inflation_cutoff=None, drop_mask=None):
    """ICE branch of the driver.

    matrix           full symmetric scipy matrix in the original bin frame (n bins)
    filter_threshold (lo, hi) modified z-score thresholds (--filterThreshold)
    chr_offsets      bin offsets of the chromosomes (len nchrom+1); needed for perchr
    perchr           statistics and correction per chromosome block (--perchr)
    extra_mask_fn    optional callable(kept_bins) -> bool mask over the bins that survived the MAD
                     filter, for the host-side --sequencedCountCutoff filter

    upper_input      ``matrix`` holds only the upper triangle (what the files store): half the upload, the mirror
                     image is built on the device (hb_csr_create_upper)
    drop_mask        uint8[bins of ``matrix``], 1 = delete the bin right after the upload (a chromosome subset chosen with
                     --chromosomes, hicCorrectMatrix.py:606-609); every other argument and every result then refers to
                     the remaining bins
    file_layout      return ``upper`` = corrected matrix as the files store it (upper triangle, zeros eliminated,
                     original n-bin frame; hicmatrix restoreMaskedBins + triu, done on the device) instead of ``matrix``

    trans_cutoff     --transCutoff (:695-699): percentile ``100 - trans_cutoff/100`` of the values stored between
                     chromosomes (hicmatrix truncTrans) -> ``max_inter``; the row sums at that point are kept for
                     ``inflation_cutoff``.  ``clip_trans`` applies the clipping truncTrans documents (values >= max_inter
                     become max_inter); off by default because the published hicmatrix statement is a comparison, not
                     an assignment, and leaves the matrix unchanged.
    inflation_cutoff --inflationCutoff (:765-775): after the correction, bins whose row sum grew by at least this
                     factor are masked (``inflated_rel``, relative to ``kept_ice``); needs ``trans_cutoff`` like the
                     reference (its ``pre_row_sum`` only exists then).

    Returns a dict:
      zero_bins      original ids removed because their row sum is 0        (:612-616)
      outliers_rel   MAD outliers, indices relative to the zero-masked frame (:656-657)
      kept           original ids of the bins that went through ICE
      bias           correction factors of the kept bins                     (:738 / :713)
      matrix         corrected CSR in the kept frame (block diagonal for perchr)
      iterations     loop bodies executed (per chromosome for perchr)
      deviation      last max|s-1| (per chromosome for perchr)
    """
    _lib.require_device()
    m = canonical_csr(matrix, dtype=None)      # values keep the file type: widened to fp64 on the device
    n = m.shape[0]
    lo, hi = filter_threshold
    with (DeviceCSR.from_upper(m, device=device) if upper_input else DeviceCSR.from_scipy(m, device=device)) as dev:
        if drop_mask is not None:
            dev.compact(np.ascontiguousarray(drop_mask, dtype=np.uint8))
            n = dev.n
        # mask all zero value bins: row sums include the diagonal and NaNs propagate, as in the
        # reference where this happens before the NaN/Inf scrub
        row_sum, _ = dev.row_stats()
        zero_mask = row_sum == 0
        zero_bins = np.flatnonzero(zero_mask)
        log.info("Removing {} zero value bins".format(len(zero_bins)))
        offsets = None if chr_offsets is None else np.asarray(chr_offsets, dtype=np.int64)
        if perchr:
            if offsets is None:
                raise ValueError("perchr needs chromosome offsets")
            dev.set_segments(offsets)
        dev.compact(zero_mask.astype(np.uint8))
        keep1 = np.flatnonzero(~zero_mask)
        dev.scrub()
        log.info("matrix contains {} data points. Sparsity {:.3f}.".format(
            dev.nnz, float(dev.nnz) / (max(dev.n, 1) ** 2)))
        if skip_diagonal:
            dev.diag_zero()
        mask, _z, _med, _mad = dev.mad_filter(lo, hi)
        outliers_rel = np.flatnonzero(mask)
        remove = mask.astype(bool)
        if extra_mask_fn is not None:
            extra = np.zeros(len(keep1), dtype=bool)
            surv = np.flatnonzero(~remove)
            extra[surv] = extra_mask_fn(keep1[surv])
            remove |= extra
        dev.compact(remove.astype(np.uint8))
        kept = keep1[~remove]
        extra_out = {}
        pre_row_sum = None
        if trans_cutoff and 0 < trans_cutoff < 100:
            if offsets is None:
                raise ValueError("trans_cutoff needs chromosome offsets")
            offsets_now = np.searchsorted(kept, offsets, side="left")       # chromosome offsets in the masked frame
            max_inter = dev.trans_percentile(offsets_now, 100 - float(trans_cutoff) / 100)
            extra_out["max_inter"] = max_inter
            if clip_trans and max_inter is not None:
                extra_out["clipped"] = dev.trans_clip(offsets_now, max_inter)
            pre_row_sum, _ = dev.row_stats()
        if inflation_cutoff and inflation_cutoff > 0 and pre_row_sum is None:
            raise ValueError("inflation_cutoff needs trans_cutoff (pre_row_sum is only defined then, "
                             "hicCorrectMatrix.py:699, 771)")
        if perchr:
            dev.keep_cis()
        if dev.symmetry_ratio() > 1e-10:
            raise ValueError("Please provide symmetric matrix!")
        dev.pack_values()      # raw counts stream as uint16 (6 B/entry instead of 12); lossless, results unchanged
        bias, iters, deviation = dev.ice(max_iter, tolerance)
        out_ids = kept
        if inflation_cutoff and inflation_cutoff > 0:
            after = dev.ice_scaled_row_sums()
            with np.errstate(divide="ignore", invalid="ignore"):
                inflated = after / pre_row_sum >= inflation_cutoff
            extra_out["inflated_rel"] = np.flatnonzero(inflated)
            extra_out["kept_ice"] = kept
            out_ids = np.where(inflated, -1, kept)      # -1: left out of the emitted matrix
            kept = kept[~inflated]
            bias = bias[~inflated]
        out = dict(zero_bins=zero_bins, outliers_rel=outliers_rel, kept=kept, bias=bias,
                   iterations=iters, deviation=deviation, n=n, **extra_out)
        if want_matrix and file_layout:
            out["upper"] = dev.ice_scaled_upper(out_ids, n)
        elif want_matrix and "inflated_rel" in extra_out:
            raise ValueError("inflation_cutoff returns the matrix in file layout only (file_layout=True)")
        elif want_matrix:
            values = dev.ice_scaled_values()
            pattern = dev.download()
            out["matrix"] = sparse.csr_matrix((values, pattern.indices, pattern.indptr), shape=pattern.shape)
    return out


def kr_pipeline(matrix, rescale=True, want_matrix=True, device=0, upper_input=False, rescale_vector=None,
                drop_mask=None, skip_diagonal=False):
    """KR branch of the driver (hicCorrectMatrix.py:742-757): no masking, no filter.
    Returns dict(x, factor, outer, matvecs, residual, upper).

    upper_input     ``matrix`` holds only the upper triangle (what the files store); the symmetric store is built on the
                    device (hb_csr_create_upper)
    rescale_vector  whether the returned ``x`` is the rescaled one (default: same as ``rescale``).  The reference's
                    --perchr + .cool branch returns the un-rescaled vector (:731) while --perchr + .h5 rescales through
                    get_normalised_matrix(True) first (:727-732)."""
    _lib.require_device()
    m = canonical_csr(matrix, dtype=None)      # values keep the file type: widened to fp64 on the device
    with (DeviceCSR.from_upper(m, device=device) if upper_input else DeviceCSR.from_scipy(m, device=device)) as dev:
        if drop_mask is not None:              # chromosome subset (--chromosomes): delete the other bins first
            dev.compact(np.ascontiguousarray(drop_mask, dtype=np.uint8))
        dev.scrub()                            # NaN / Inf -> 0 happens before the method switch (:624-625)
        if skip_diagonal:
            dev.diag_zero()                    # so does --skipDiagonal (:648-649)
        dev.pack_values()
        x, outer, matvecs, residual = dev.kr()
        factor = 1.0
        x_raw = x
        if rescale:
            x, factor = dev.kr_rescale(x, KR_DIAG_EPS)
        out = dict(x=x if (rescale if rescale_vector is None else rescale_vector) else x_raw, factor=factor, outer=outer,
                   matvecs=matvecs, residual=residual)
        if want_matrix:
            out["upper"] = dev.kr_scaled_upper(x, KR_DIAG_EPS)
    return out


def scatter_to_frame(corrected_kept, kept, n):
    """Re-insert the masked bins as empty rows/columns (hicmatrix restoreMaskedBins)."""
    coo = corrected_kept.tocoo()
    kept = np.asarray(kept)
    return sparse.csr_matrix((coo.data, (kept[coo.row], kept[coo.col])), shape=(n, n))


# ---------------------------------------------------------------------------------------------------------
# Several GPUs: one process per GPU (torchrun), nnz-balanced row blocks, the same stages on every block.
# Every rank holds the host matrix (it came from the same file) and uploads only its rows; rank 0 returns
# the assembled result, the other ranks return None.
# ---------------------------------------------------------------------------------------------------------
def _dist_env():
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(), dist.get_world_size()


def _gather_rows_to_root(values, dist, rank, world, device):
    """concatenate per-rank 1-D float64 arrays on rank 0 (row blocks are contiguous and ordered by rank)."""
    import torch
    sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
    sizes[rank] = len(values)
    dist.all_reduce(sizes)
    sizes = sizes.cpu().numpy()
    mine = torch.from_numpy(np.ascontiguousarray(values)).cuda()
    if rank == 0:
        parts = [values]
        for src in range(1, world):
            buf = torch.empty(int(sizes[src]), dtype=mine.dtype, device="cuda")
            if sizes[src]:
                dist.recv(buf, src=src)
            parts.append(buf.cpu().numpy())
        return np.concatenate(parts)
    if len(values):
        dist.send(mine, dst=0)
    return None


def _gather_upper_to_root(upper, dist, rank, world):
    """Row blocks of the file-layout matrix (every rank holds all n rows, only its own non-empty) -> one CSR on rank 0.
    Blocks are contiguous and ordered by rank, so concatenating indices / data in rank order is row-major order."""
    row_counts = np.diff(upper.indptr).astype(np.float64)
    idx = _gather_rows_to_root(upper.indices.astype(np.float64), dist, rank, world, None)
    val = _gather_rows_to_root(upper.data, dist, rank, world, None)
    import torch
    counts = torch.from_numpy(row_counts).cuda()
    dist.all_reduce(counts)                                  # x + 0: each row is owned by exactly one rank
    if rank != 0:
        return None
    indptr = np.concatenate([[0], np.cumsum(counts.cpu().numpy().astype(np.int64))])
    return sparse.csr_matrix((val, idx.astype(np.int32), indptr), shape=upper.shape)


def _open_row_block(m, upper_input, drop_mask, rank, world, device):
    """this rank's nnz-balanced row block of the symmetric matrix on its GPU -> (DeviceCSR, (a, b), n).
    ``upper_input``: every rank uploads the file's upper triangle, builds the symmetric store in HBM, deletes the bins of
    ``drop_mask`` and keeps its own rows (hb_csr_keep_rows): the symmetric matrix never exists on the host."""
    from .dist import nnz_balanced_partition
    if upper_input:
        dev = DeviceCSR.from_upper(m, device=device)
        if drop_mask is not None:
            dev.compact(np.ascontiguousarray(drop_mask, dtype=np.uint8))
        n = dev.n
        bounds = nnz_balanced_partition(dev.download_indptr(), world)
        a, b = int(bounds[rank]), int(bounds[rank + 1])
        dev.keep_rows(a, b)
        return dev, (a, b), n
    n = m.shape[0]
    bounds = nnz_balanced_partition(m.indptr, world)
    a, b = int(bounds[rank]), int(bounds[rank + 1])
    return DeviceCSR.from_scipy(m, device=device, row_range=(a, b)), (a, b), n


def _attach_exchange(dev, rank, world):
    """Register the host collective (row sums of the filter, symmetry digests, KR rescale sums) and, where the GPUs can
    reach each other (one NVLink / P2P domain, <= 8 ranks), switch the per-iteration exchange of ICE / KR to peer stores
    from the SpMV epilogue: the solvers then run as ONE persistent launch per rank with no host or NCCL call inside.
    Same results bit for bit either way.  HB_EXCHANGE=nccl keeps the host collective for everything."""
    import os

    from .dist import attach_comm, attach_peer
    keep = attach_comm(dev, rank, world)
    if os.environ.get("HB_EXCHANGE", "peer") != "nccl" and world <= 8:
        attach_peer(dev, rank, world)
    return keep


def _assemble_rows(dev, local_values, dist):
    """row-block vector -> the whole n-vector on every rank (x + 0 over the ranks: exact)"""
    import torch
    n_rows, n_cols, row0, _nnz = dev.dims()
    full = torch.zeros(n_cols, dtype=torch.float64, device="cuda")
    if n_rows:
        full[row0:row0 + n_rows] = torch.from_numpy(np.ascontiguousarray(local_values, dtype=np.float64)).cuda()
    dist.all_reduce(full)
    return full.cpu().numpy()


def ice_pipeline_sharded(matrix, filter_threshold, max_iter=500, tolerance=1e-5, chr_offsets=None, perchr=False,
                         skip_diagonal=False, want_matrix=True, file_layout=False, upper_input=False, drop_mask=None,
                         extra_mask_fn=None, trans_cutoff=None, clip_trans=False, inflation_cutoff=None):
    """ice_pipeline on row blocks across the initialised torch.distributed group (NCCL).  The symmetry test is the
    exact-symmetry digest screen summed over the blocks (one tiny allreduce); only a matrix that is not exactly symmetric
    pays for the reference's ratio (iterativeCorrection.py:35): one all-to-all that brings every entry to the owner of its
    transposed partner (DeviceCSR.symmetry_ratio_sharded).

    ``file_layout``: every rank emits its rows of the corrected matrix already in file layout (upper triangle, zeros
    eliminated, original frame: hb_ice_scaled_upper works on a row block) and rank 0 concatenates the blocks -> ``upper``;
    otherwise rank 0 rebuilds the kept-frame pattern on the host -> ``matrix``."""
    import torch

    from .dist import attach_comm, nnz_balanced_partition
    dist, rank, world = _dist_env()
    if dist is None or world == 1:
        return ice_pipeline(matrix, filter_threshold, max_iter, tolerance, chr_offsets, perchr, skip_diagonal,
                            device=torch.cuda.current_device(), want_matrix=want_matrix, file_layout=file_layout,
                            upper_input=upper_input, drop_mask=drop_mask, extra_mask_fn=extra_mask_fn,
                            trans_cutoff=trans_cutoff, clip_trans=clip_trans, inflation_cutoff=inflation_cutoff)
    device = torch.cuda.current_device()
    m = canonical_csr(matrix, dtype=None)      # values keep the file type: widened to fp64 on the device
    lo, hi = filter_threshold
    if upper_input and want_matrix and not file_layout:
        raise ValueError("with upper_input the corrected matrix is returned in file layout only (file_layout=True)")
    upper_local = None
    dev, (a, b), n = _open_row_block(m, upper_input, drop_mask, rank, world, device)
    with dev:
        keep_alive = _attach_exchange(dev, rank, world)
        rs_local, _ = dev.row_stats()
        rs = torch.zeros(n, dtype=torch.float64, device="cuda")
        rs[a:b] = torch.from_numpy(rs_local).cuda()
        dist.all_reduce(rs)                                   # x + 0: exact assembly of the row sums
        zero_mask = (rs == 0).cpu().numpy()
        zero_bins = np.flatnonzero(zero_mask)
        if perchr:
            dev.set_segments(np.asarray(chr_offsets, dtype=np.int64))
        dev.compact(zero_mask.astype(np.uint8))
        keep1 = np.flatnonzero(~zero_mask)
        dev.scrub()
        if skip_diagonal:
            dev.diag_zero()
        mask, _z, _med, _mad = dev.mad_filter(lo, hi)         # one collective inside, mask replicated
        outliers_rel = np.flatnonzero(mask)
        remove = mask.astype(bool)
        if extra_mask_fn is not None:                         # --sequencedCountCutoff: O(n) host work, replicated
            extra = np.zeros(len(keep1), dtype=bool)
            surv = np.flatnonzero(~remove)
            extra[surv] = extra_mask_fn(keep1[surv])
            remove |= extra
        dev.compact(remove.astype(np.uint8))
        kept = keep1[~remove]
        extra_out = {}
        pre_row_sum = None
        offsets = None if chr_offsets is None else np.asarray(chr_offsets, dtype=np.int64)
        if trans_cutoff and 0 < trans_cutoff < 100:
            if offsets is None:
                raise ValueError("trans_cutoff needs chromosome offsets")
            offsets_now = np.searchsorted(kept, offsets, side="left")
            # 8 histogram passes of the radix select, each summed over the row blocks by one small allreduce
            max_inter = dev.trans_percentile(offsets_now, 100 - float(trans_cutoff) / 100)
            extra_out["max_inter"] = max_inter
            if clip_trans and max_inter is not None:
                clipped = torch.tensor([float(dev.trans_clip(offsets_now, max_inter))], dtype=torch.float64, device="cuda")
                dist.all_reduce(clipped)
                extra_out["clipped"] = int(clipped.item())
            pre_row_sum = _assemble_rows(dev, dev.row_stats()[0], dist)
        if inflation_cutoff and inflation_cutoff > 0 and pre_row_sum is None:
            raise ValueError("inflation_cutoff needs trans_cutoff (pre_row_sum is only defined then, "
                             "hicCorrectMatrix.py:699, 771)")
        if perchr:
            dev.keep_cis()
        if dev.symmetry_ratio_sharded(dist, rank, world) > 1e-10:   # exact screen; exact ratio via a block exchange if needed
            raise ValueError("Please provide symmetric matrix!")
        dev.pack_values()
        bias, iters, deviation = dev.ice(max_iter, tolerance)  # one exchange per iteration
        out_ids = kept
        if inflation_cutoff and inflation_cutoff > 0:
            if want_matrix and not file_layout:
                raise ValueError("inflation_cutoff returns the matrix in file layout only (file_layout=True)")
            after = _assemble_rows(dev, dev.ice_scaled_row_sums(), dist)
            with np.errstate(divide="ignore", invalid="ignore"):
                inflated = after / pre_row_sum >= inflation_cutoff
            extra_out["inflated_rel"] = np.flatnonzero(inflated)
            extra_out["kept_ice"] = kept
            out_ids = np.where(inflated, -1, kept)
            kept = kept[~inflated]
            bias = bias[~inflated]
        values = None
        if want_matrix and file_layout:
            upper_local = dev.ice_scaled_upper(out_ids, n)
        elif want_matrix:
            values = dev.ice_scaled_values()
        del keep_alive
    out = dict(zero_bins=zero_bins, outliers_rel=outliers_rel, kept=kept, bias=bias, iterations=iters,
               deviation=deviation, n=n, **extra_out)
    if want_matrix and file_layout:
        upper = _gather_upper_to_root(upper_local, dist, rank, world)
        if rank == 0:
            out["upper"] = upper
    elif want_matrix:
        all_values = _gather_rows_to_root(values, dist, rank, world, device)
        if rank == 0:
            pattern = sparse.csr_matrix(m[kept, :][:, kept])
            pattern.sort_indices()
            if perchr:
                seg = np.searchsorted(np.asarray(chr_offsets), kept, side="right")
                coo = pattern.tocoo()
                cis = seg[coo.row] == seg[coo.col]
                pattern = sparse.csr_matrix((coo.data[cis], (coo.row[cis], coo.col[cis])), shape=pattern.shape)
                pattern.sort_indices()
            if skip_diagonal:
                pattern.setdiag(0)
                pattern.eliminate_zeros()
            assert pattern.nnz == len(all_values), "device and host patterns disagree"
            out["matrix"] = sparse.csr_matrix((all_values, pattern.indices, pattern.indptr), shape=pattern.shape)
    return out if rank == 0 else None


def kr_pipeline_sharded(matrix, rescale=True, upper_input=False, drop_mask=None, skip_diagonal=False, want_matrix=False):
    """kr_pipeline on row blocks: x is replicated on every rank; rank 0 also gets the rescale factor.
    ``want_matrix``: every rank emits its rows of the balanced matrix in file layout (hb_kr_scaled_upper works on a row
    block) and rank 0 concatenates them -> ``upper`` (what a genome-wide KR run to .h5 stores)."""
    import torch

    from .dist import attach_comm, nnz_balanced_partition
    dist, rank, world = _dist_env()
    if dist is None or world == 1:
        return kr_pipeline(matrix, rescale=rescale, want_matrix=want_matrix, device=torch.cuda.current_device(),
                           upper_input=upper_input, drop_mask=drop_mask, skip_diagonal=skip_diagonal)
    device = torch.cuda.current_device()
    m = canonical_csr(matrix, dtype=None)      # values keep the file type: widened to fp64 on the device
    dev, _block, _n = _open_row_block(m, upper_input, drop_mask, rank, world, device)
    with dev:
        keep_alive = _attach_exchange(dev, rank, world)
        dev.scrub()
        if skip_diagonal:
            dev.diag_zero()
        dev.pack_values()
        x, outer, matvecs, residual = dev.kr()
        factor = 1.0
        if rescale:
            x, factor = dev.kr_rescale(x, KR_DIAG_EPS)
        upper = None
        if want_matrix:
            n_rows, n_cols, row0, _nnz = dev.dims()
            block = dev.kr_scaled_upper(x, KR_DIAG_EPS)           # the block's rows, global column ids
            indptr = np.zeros(n_cols + 1, dtype=np.int64)
            indptr[row0 + 1:row0 + n_rows + 1] = block.indptr[1:]
            indptr[row0 + n_rows + 1:] = block.indptr[-1]
            upper = _gather_upper_to_root(sparse.csr_matrix((block.data, block.indices, indptr), shape=(n_cols, n_cols)),
                                          dist, rank, world)
        del keep_alive
    out = dict(x=x, factor=factor, outer=outer, matvecs=matvecs, residual=residual)
    if want_matrix and rank == 0:
        out["upper"] = upper
    return out
