/*
 * libhicbalance -- B200-native (sm_100a) Hi-C matrix balancing: ICE iterative
 * correction, Knight-Ruiz balancing and the MAD bin filter on a device-resident
 * symmetric CSR store.
 *
 * This header is the drop-in boundary of the hot path.  The reference
 * (HiCExplorer 3.7.6) is pure Python and reaches native code for this path in
 * exactly one place, the pybind11 object `krbalancing.kr_balancing`
 * (hicexplorer/hicCorrectMatrix.py:17, 718-732, 744-757); its ICE is numpy
 * (hicexplorer/iterativeCorrection.py:10-86).  Every entry point below cites
 * the reference interface it replaces.  INTEGRATION.md shows the ctypes stubs
 * a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain C ABI: pointers + sizes, no C++/torch types.
 *   - HOST API  (hb_*):      array arguments are caller-owned host buffers; the
 *     call is synchronous (returns after the device work finished), like the
 *     reference's single-threaded functions.
 *   - DEVICE API (hb_dev_*): array arguments are device pointers owned by the
 *     caller (e.g. torch tensors); work is enqueued on `stream` (a cudaStream_t
 *     passed as void*, NULL = the handle's own stream) and is asynchronous.
 *     This is what the multi-GPU driver and the HBM-resident benchmark use.
 *   - a handle owns one CUDA stream and is not thread-safe.
 *   - every function returns an hb_status; hb_last_error() gives the message of
 *     the last failure on the calling thread.
 *   - matrix layout in HBM: indptr int64[n_rows+1] (block-local offsets),
 *     indices int32[nnz] (GLOBAL column ids, ascending inside a row),
 *     data fp64[nnz]; FULL symmetric storage (both triangles), which is what
 *     the reference holds in memory.
 */
#ifndef HICBALANCE_H
#define HICBALANCE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 1

#if defined(__GNUC__)
#define HB_API __attribute__((visibility("default")))
#else
#define HB_API
#endif

typedef enum hb_status {
    HB_OK = 0,
    /* iterativeCorrection.py:35-36  -> ValueError("Please provide symmetric matrix!") */
    HB_ERR_NOT_SYMMETRIC = 1,
    /* iterativeCorrection.py:58-62  -> log.error + exit(1): a scaled value exceeded 1e100 */
    HB_ERR_OVERFLOW_ITER = 2,
    /* iterativeCorrection.py:80-84  -> log.error + exit(1): a final value exceeded 1e10 */
    HB_ERR_OVERFLOW_FINAL = 3,
    /* krbalancing asserts (rho > 0, finite, min(x) > 0) */
    HB_ERR_KR_BREAKDOWN = 4,
    HB_ERR_INVALID_ARG = 5,
    HB_ERR_NO_DEVICE = 6,
    HB_ERR_CUDA = 100
} hb_status;

typedef struct hb_csr hb_csr; /* opaque: device CSR block + work vectors + stream */

/* ---- library ------------------------------------------------------------ */
HB_API int hb_abi_version(void);
HB_API const char* hb_last_error(void);
/* number of visible CUDA devices (0 if none / driver missing); never fails */
HB_API int hb_device_count(void);
/* pinned host memory for callers that want full-rate H2D/D2H */
HB_API int hb_pinned_alloc(void** out, uint64_t bytes);
HB_API int hb_pinned_free(void* p);
/* number of kernel launches issued by this library in this process (bench.py: gpu_launches) */
HB_API uint64_t hb_launch_count(void);

/* ---- device CSR store ------------------------------------------------------
 * Replaces: the scipy CSR the reference keeps in `ma.matrix`
 * (hicCorrectMatrix.py:603-626) and krbalancing's constructor copy of
 * (indptr, indices, data) (hicCorrectMatrix.py:744-747).
 *
 * hb_csr_create: copy a host CSR block to `device`.  The block holds rows
 * [row0, row0 + n_rows) of an n_cols x n_cols symmetric matrix (single GPU:
 * row0 = 0, n_rows = n_cols).  indptr may start at any offset; it is rebased.
 * indices_i64 != 0 means `indices` points at int64 (krbalancing's ctor
 * signature) instead of int32.  The block is validated on the device before the call returns (one pass over the column
 * ids): a non-monotone indptr, a column id outside [0, n_cols) or columns that are not strictly ascending inside a row
 * give HB_ERR_INVALID_ARG -- no later kernel ever gathers through an unchecked index. */
HB_API int hb_csr_create(hb_csr** out, int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz,
                  const int64_t* indptr, const void* indices, int indices_i64, const double* data,
                  int device);
/* The same two constructors for values in the file's own type (counts are int32 / int64 in .cool files, float32 after
 * hicNormalize): the values are widened to the fp64 store on the device, so the host needs no fp64 copy and fewer bytes
 * cross PCIe.  value_kind: */
enum { HB_VALUES_F64 = 0, HB_VALUES_F32 = 1, HB_VALUES_I32 = 2, HB_VALUES_I64 = 3 };
HB_API int hb_csr_create_typed(hb_csr** out, int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz,
                        const int64_t* indptr, const void* indices, int indices_i64, const void* data, int value_kind,
                        int device);
HB_API int hb_csr_create_upper_typed(hb_csr** out, int64_t n, int64_t nnz_upper, const int64_t* indptr,
                              const int32_t* indices, const void* data, int value_kind, int device);
/* Same store from the UPPER triangle only (what .h5 / .cool files hold; hicmatrix symmetrises on load,
 * call site hicCorrectMatrix.py:603-609): half the bytes cross PCIe, the mirror image is built in HBM with ascending
 * columns (deterministic layout, identical to a host-side fill).  Entries below the diagonal are an error. */
HB_API int hb_csr_create_upper(hb_csr** out, int64_t n, int64_t nnz_upper, const int64_t* indptr, const int32_t* indices,
                        const double* data, int device);
/* The device half of hb_csr_create_upper on a handle that already holds an upper-triangular matrix (e.g. after
 * hb_csr_keep_upper): the store becomes the full symmetric matrix M + triu(M, 1)^T. */
HB_API int hb_csr_symmetrize_upper(hb_csr* h);
/* New store C with C[i + dr][j + dc] = A[i][j] for the targets inside the frame: one term of the running-window merge
 * (hicMergeMatrixBins.py:88-186 sums (2h+1)^2 shifted copies of the upper triangle). */
HB_API int hb_csr_shifted_copy(hb_csr** out, const hb_csr* src, int64_t dr, int64_t dc);
/* wrap device arrays already in HBM (not copied, not owned). indptr must be block-local (indptr[0]==0). */
HB_API int hb_dev_csr_wrap(hb_csr** out, int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz,
                    const int64_t* d_indptr, const int32_t* d_indices, const double* d_data, int device);
HB_API int hb_csr_destroy(hb_csr* h);
HB_API int hb_csr_shape(const hb_csr* h, int64_t* n_rows, int64_t* n_cols, int64_t* row0, int64_t* nnz);
/* copy the (possibly compacted) block back to host buffers sized from hb_csr_shape; any pointer may be NULL */
HB_API int hb_csr_download(hb_csr* h, int64_t* indptr, int32_t* indices, double* data);
/* device pointers of the store (for zero-copy torch views) */
HB_API int hb_dev_csr_pointers(hb_csr* h, const int64_t** d_indptr, const int32_t** d_indices, const double** d_data);
HB_API int hb_sync(hb_csr* h);

/* Independent diagonal blocks (`--perchr`, hicCorrectMatrix.py:703-734).
 * seg_offsets[nseg+1] are ascending GLOBAL row offsets with seg_offsets[0]=0 and
 * seg_offsets[nseg]=n_cols.  Every later ICE / KR / filter call treats each
 * segment as its own matrix (own mean, own convergence, own medians).
 * Entries whose row and column lie in different segments must be removed first
 * with hb_csr_keep_cis (the reference drops them: hicCorrectMatrix.py:702,711). */
HB_API int hb_set_segments(hb_csr* h, int nseg, const int64_t* seg_offsets);
HB_API int hb_csr_keep_cis(hb_csr* h);

/* Sharded runs (one process per GPU, nnz-balanced row blocks).  The path has exactly
 * one exchange step per SpMV: every rank writes the rows it owns into a zeroed fp64 vector and the
 * vectors are summed across ranks (x + 0 is exact, so the result is bit-identical to a gather).
 * The library calls `fn` for it: fn must sum-allreduce d_buf[0..count) over the ranks, ordered after
 * the work already enqueued on `stream` and complete (stream-ordered) when it returns.  The Python
 * host passes a torch.distributed (NCCL) all_reduce here.  With comm set, hb_ice_run / hb_kr_run
 * accept a row block instead of the full matrix; the O(n) vector work runs replicated on every rank. */
/* Turn a handle that holds the FULL matrix into the row block [row_begin, row_end) of it (global column ids kept), as if
 * it had been created with row0 = row_begin: how a rank of a sharded run gets its block when the whole upper triangle
 * was uploaded and symmetrised on the device (hb_csr_create_upper) -- no symmetric matrix on the host. */
HB_API int hb_csr_keep_rows(hb_csr* h, int64_t row_begin, int64_t row_end);
typedef int (*hb_allreduce_fn)(void* ctx, void* d_buf, int64_t count, void* stream);
HB_API int hb_set_comm(hb_csr* h, int rank, int world, hb_allreduce_fn fn, void* ctx);

/* Peer-store exchange (alternative to the callback; one process per GPU inside one NVLink domain, world <= 8).
 * The SpMV epilogue stores every row's result directly into all ranks' replicas of the exchange vector (P2P stores
 * through CUDA IPC mappings), then publishes an arrival flag; no collective is launched in the loop.
 *   hb_peer_export: allocate this rank's symmetric region and return its 64-byte cudaIpcMemHandle_t.
 *   hb_peer_attach: all_handles = world x 64 bytes in rank order (gathered by the host with any collective);
 *                   maps the peers' regions.  The host must barrier once after every rank has attached.
 * Afterwards hb_ice_run / hb_kr_run / hb_dev_ice_spmv / hb_dev_ice_update use the region (pass d_exch = NULL) and the
 * host must NOT all-reduce.  hb_peer_active() tells which mode a handle is in. */
HB_API int hb_peer_export(hb_csr* h, int rank, int world, void* ipc_handle_out);
HB_API int hb_peer_attach(hb_csr* h, const void* all_handles);
/* Same exchange when all ranks live in ONE process (one host thread + handle per rank; several GPUs with peer access,
 * or several row blocks on one GPU): ranks[world] = every rank's handle after its hb_peer_export, in rank order.
 * Ranks that share a GPU must fit on it together: their persistent kernels wait for each other. */
HB_API int hb_peer_attach_inprocess(hb_csr* h, hb_csr* const* ranks);
HB_API int hb_peer_active(const hb_csr* h);

/* ---- pre-filter (device) ---------------------------------------------------
 * hb_scrub: NaN -> 0 and +-Inf -> 0 in the stored values
 *   (utilities.py:111-128 via hicCorrectMatrix.py:624-625; iterativeCorrection.py:28-30).
 *   counts[0] = #NaN, counts[1] = #Inf replaced.
 * hb_row_stats: row_sum[i] (diagonal included, hicCorrectMatrix.py:613) and
 *   diag[i] (hicCorrectMatrix.py:568,588) for the block's rows; either may be NULL.
 * hb_mad_filter: modified z-score of points = row_sum - diag and the outlier
 *   mask (class MAD, hicCorrectMatrix.py:349-391; filter_by_zscore :548-592;
 *   per segment when segments are set).  On a row block (hb_set_comm) the row statistics are assembled with one
 *   sum-allreduce and the O(n) selection runs replicated: every rank gets the same full-length mask.
 *   mask_out[i] = 1 if z < lo or z > hi; zscore_out optional (NULL).
 * hb_csr_compact: delete the rows AND columns with remove_mask[i] != 0 and
 *   renumber (hicmatrix maskBins as used at hicCorrectMatrix.py:616, 672).
 *   Segment offsets are renumbered with it.  On a row block every rank passes the same global mask and keeps
 *   its own surviving rows (row0 / n_rows / n_cols are renumbered).
 * hb_diag_zero: --skipDiagonal (hicCorrectMatrix.py:648-649).
 * hb_symmetry: the reference's criterion (iterativeCorrection.py:35):
 *   ratio = mean|M - M^T| / |mean M|; the caller raises when ratio > 1e-10.  On a row block (hb_set_comm) the
 *   exact-symmetry digests of the blocks are summed with one 16-double allreduce: ratio 0 when the whole matrix equals
 *   its transpose exactly, else HB_ERR_NOT_SYMMETRIC with ratio 1 (the exact ratio needs other ranks' entries). */
HB_API int hb_scrub(hb_csr* h, int64_t counts[2]);
/* hb_row_stats restricted to each row's own segment (chromosome block): what `diagnostic_plot --perchr` histograms
 * (hicCorrectMatrix.py:503-507) and what the per-chromosome MAD filter works on (:557-568). */
HB_API int hb_row_stats_cis(hb_csr* h, double* row_sum, double* diag);
/* Optional: build a LOSSLESS narrow copy of the stored values for the streaming kernels (ICE / KR mat-vecs):
 * uint16 when every value is an integer count <= 65535 (kind 1, 6 B per entry streamed instead of 12), float when
 * every value is exactly representable in fp32 (kind 2, 8 B), none otherwise (kind 0).  Values are widened to fp64
 * before the first multiply, so every result is bit-identical to the fp64 layout.  The fp64 array stays the store of
 * record (filter, emit, download); the copy is dropped whenever the values change (scrub, compaction). */
HB_API int hb_csr_pack_values(hb_csr* h, int* kind_out);
/* hb_csr_create does the same on the way in when it can: fp64 values that are all integer counts <= 65535 are
 * converted to uint16 by host threads while still in host memory (bit-exact round trip or refused), 2 bytes per entry
 * cross PCIe instead of 8, and the device writes both the uint16 copy (kind 1) and the fp64 store from them.  With a
 * pinned source a feeder thread sends chunks from the other end of the array as they are whenever the link would
 * otherwise wait for the converters (few cores per rank), and that part is narrowed on the device.
 * HB_NARROW_INGEST=0 switches the whole mechanism off; HB_INGEST_DIRECT=1 / 0 forces the feeder on / off (default: on
 * when the rank converts with at most 8 host threads).  hb_csr_ingest_info: the layout a
 * store currently has and the bytes its creation moved from host to device. */
HB_API int hb_csr_ingest_info(const hb_csr* h, int* pack_kind, int64_t* h2d_bytes);
HB_API int hb_row_stats(hb_csr* h, double* row_sum, double* diag);
HB_API int hb_mad_filter(hb_csr* h, double lo, double hi, uint8_t* mask_out, double* zscore_out,
                  double* median_out /*nseg*/, double* mad_out /*nseg*/);
HB_API int hb_csr_compact(hb_csr* h, const uint8_t* remove_mask);
HB_API int hb_diag_zero(hb_csr* h);
/* keep only the upper triangle (column >= row) of the block: the file layout (inverse of hb_csr_create_upper) */
HB_API int hb_csr_keep_upper(hb_csr* h);
HB_API int hb_symmetry(hb_csr* h, double* ratio_out);
/* Exact sum |M - M^T| on row blocks (the numerator of iterativeCorrection.py:35 when hb_symmetry's exact-symmetry
 * screen says "not identical" on a sharded matrix).  hb_dev_extract_cols: the block's entries whose column lies in
 * [col_lo, col_hi) as DEVICE triplets (global row, column, value), any order; call once with NULL outputs for the
 * count.  The host sends them to the rank that owns those rows; hb_dev_partner_sums looks every received entry
 * a_ij up in the local row j and returns sum of |a_ij - a_ji| (2 |a_ij| where a_ji is not stored).  Summed over the
 * ranks, divided by sum M: the reference's ratio. */
HB_API int hb_dev_extract_cols(hb_csr* h, int64_t col_lo, int64_t col_hi, int64_t* count_out, int32_t* d_rows, int32_t* d_cols,
                               double* d_vals);
HB_API int hb_dev_partner_sums(hb_csr* h, int64_t n, const int32_t* d_rows, const int32_t* d_cols, const double* d_vals,
                               double* sum_out);

/* ---- bin merging (the step before correction) ----------------------------------
 * Replaces reduce_matrix(matrix, bins_to_merge, use_triu=True, diagonal=True) (hicexplorer/reduceMatrix.py:12-232)
 * as called by hicMergeMatrixBins (hicexplorer/hicMergeMatrixBins.py:264).  bin_map[n_cols]: merged bin id of every
 * bin, -1 = dropped; ids must appear in ascending order 0..n_new-1, each covering <= 32 CONSECUTIVE bins (what
 * --numBins <= 32 produces).  The store becomes the n_new x n_new full symmetric merged matrix (exact zeros
 * eliminated, columns ascending); segments are reset.  Needs the full matrix on one device. */
HB_API int hb_csr_merge_bins(hb_csr* h, const int32_t* bin_map, int64_t n_new);

/* ---- ICE -------------------------------------------------------------------
 * Replaces iterativeCorrection(matrix, v, M, tolerance, verbose)
 * (hicexplorer/iterativeCorrection.py:10-86).
 *
 * hb_ice_run: iterate until max|s-1| < tol (strict, tested after the rescale of
 *   that iteration) or max_iter passes.  bias_out[n_cols] is the divisive bias,
 *   already normalised to mean 1 over its non-zero entries (lines 77-78).
 *   iters_out[nseg]: loop bodies executed per segment (the reference LOGS this
 *   value + 1).  dev_out[nseg]: last deviation.  Returns HB_ERR_OVERFLOW_ITER
 *   when a rescaled value exceeded 1e100 (line 58).
 * hb_ice_scaled_values: data_out[nnz] = A_ij / (b_i b_j) in the store's entry
 *   order (lines 56-57, 79); returns HB_ERR_OVERFLOW_ITER / _FINAL per lines
 *   58 / 80 for the final matrix.  max_out optional.
 * hb_ice_run needs the full matrix on one device unless hb_set_comm was called. */
HB_API int hb_ice_run(hb_csr* h, int max_iter, double tol, double* bias_out, int32_t* iters_out,
               double* dev_out);
HB_API int hb_ice_scaled_values(hb_csr* h, double* data_out, double* max_out);
/* The corrected matrix in FILE layout: upper triangle, exact zeros eliminated, scattered back into the original
 * n_orig-bin frame (orig_ids[i] = original id of bin i, strictly ascending; -1 = leave the bin out, which is how
 * --inflationCutoff masks bins after the correction, hicCorrectMatrix.py:774) -- hicmatrix restoreMaskedBins + triu at
 * save time (hicCorrectMatrix.py:779), fused with the final scaling.  Two calls: with indptr/indices/data NULL to get
 * nnz_out, then with indptr[n_orig+1], indices[nnz], data[nnz].  Same overflow codes as hb_ice_scaled_values. */
HB_API int hb_ice_scaled_upper(hb_csr* h, const int64_t* orig_ids, int64_t n_orig, int64_t* nnz_out, int64_t* indptr,
                        int32_t* indices, double* data);

/* ---- element-wise steps of the workflow around the correction ----------------------
 * hicNormalize (hicexplorer/hicNormalize.py:63-153), on the resident store, in the float32 arithmetic the reference
 * computes in (`matrix.data.astype(np.float32)`):
 *   hb_csr_value_stats: sum of the stored values (fp64; `hic_ma.matrix.sum()`, :70) and the float32 minimum / maximum of
 *     the stored values (:82-83), after NaN/Inf -> 0 when scrub != 0 (:77-81).  min/max are NaN for an empty store.
 *   hb_csr_normalize: value <- float32 op, then NaN/Inf -> 0, then value < threshold -> 0 (:90-96).
 *     mode 0 norm_range  (f - a) / b   a = min, b = float32(max) - float32(min)          (:84-88)
 *     mode 1 smallest    f / a         a = sum_i / sum_smallest                           (:110-111)
 *     mode 2 multiplicative f * a      a = --multiplicativeValue                          (:137)
 *     mode 3 cast only (the matrix with the smallest sum, :105-106)
 *     scrub_first != 0: NaN/Inf -> 0 before the operation as well (:77-81, :107-109, :131-135).
 *   hb_csr_eliminate_zeros: drop stored entries that are exactly 0 (`matrix.eliminate_zeros()`, :93, :97).
 * hicSumMatrices (hicexplorer/hicSumMatrices.py:59):
 *   hb_csr_add: h <- h + other (same shape, same device): union of the two sorted rows, values added, cells that sum to
 *     exactly 0 dropped like scipy's csr + csr.  `other` is not modified. */
HB_API int hb_csr_value_stats(hb_csr* h, int scrub, double* sum_out, double* min_out, double* max_out);
HB_API int hb_csr_normalize(hb_csr* h, int mode, double a, double b, double threshold, int scrub_first);
HB_API int hb_csr_eliminate_zeros(hb_csr* h);
HB_API int hb_csr_add(hb_csr* h, const hb_csr* other);

/* ---- observed / expected (hicTransform) ------------------------------------------------
 * Replaces the per-entry work of hicexplorer/utilities.py:293-338, 356-444, 479-600 as driven by hicTransform.py:162-215.
 * Segments (hb_set_segments + hb_csr_keep_cis) give --perChromosome: slot seg_off[s] + d of the n-vectors belongs to
 * distance d of chromosome s.
 *   hb_distance_stats: sums_out[.] = sum, counts_out[.] = number (may be NULL) of the stored NON-ZERO values at
 *     genomic distance d = |row - col|.
 *   hb_obs_exp_apply: value <- float32(value) / expected[index(d)] computed in fp64 like numpy does for
 *     float32_array / float64_array; index_mode 0: d (obs_exp_non_zero), 1: ceil(d / 2) (obs_exp, obs_exp_lieberman --
 *     sic, utilities.py:489, 581); NaN/Inf -> 0; out_mode 0: keep fp64, 1: truncate to an integer (`.astype(data_type)`
 *     of an integer input), 2: round to float32.  row_sum[n] + seg_total[nseg] (both or neither): the ligation factor
 *     expected *= row_sum[row] * row_sum[col] / total (utilities.py:533-534).  Zeros are left for
 *     hb_csr_eliminate_zeros. */
HB_API int hb_distance_stats(hb_csr* h, double* sums_out, int64_t* counts_out);
HB_API int hb_obs_exp_apply(hb_csr* h, const double* expected, int index_mode, int out_mode, const double* row_sum,
                     const double* seg_total);

/* ---- covariance / Pearson correlation (hicTransform --method covariance | pearson) -----------------------
 * Replaces np.cov(submatrix.todense()) / np.corrcoef(...) + convertNansToZeros + convertInfsToZeros + csr_matrix(...)
 * (hicexplorer/hicTransform.py:105-113), per chromosome block [lo, hi) (hicTransform.py:213-248) or on the whole matrix
 * (lo = 0, hi = n).  Rows of the block are the variables, its columns the observations, ddof = 1; pearson != 0 divides by
 * the standard deviations and clips to [-1, 1] like np.corrcoef.  The block is densified on the device (16 m^2 bytes)
 * and the product is one fp64 cuBLAS DGEMM (libcublas is opened at run time).  Two calls: with indptr = indices = data
 * = NULL the block is computed and kept, *nnz_out = number of non-zero cells; the second call (same lo, hi) fills
 * indptr[m+1] (block-local rows), indices (GLOBAL columns, ascending) and data, and releases the block.  Full matrix on
 * one device. */
HB_API int hb_block_correlation(hb_csr* h, int64_t lo, int64_t hi, int pearson, int64_t* nnz_out, int64_t* indptr,
                         int32_t* indices, double* data);

/* ---- optional ICE-only steps around the correction -----------------------------
 * --transCutoff (hicCorrectMatrix.py:695-699 -> hicmatrix truncTrans): trans entries are those whose row and column lie
 * in different chromosomes (chr_offsets[nchr+1], ascending, in the CURRENT bin frame).
 *   hb_trans_order_stats: n_trans_out = number of stored trans entries; when v_k / v_k1 are given, the k-th and
 *     (k+1)-th smallest trans values (0-based, exact: radix select on the fp64 bit patterns; v_k1 = v_k at the top) --
 *     the two order statistics np.percentile interpolates between.  Call once with v_k = v_k1 = NULL for the count.
 *   hb_trans_clip: every trans value >= max_inter becomes max_inter (the clipping truncTrans documents).
 * --inflationCutoff (hicCorrectMatrix.py:765-775):
 *   hb_ice_scaled_row_sums: out[i] = sum_j W_ij of the corrected matrix for the block's rows, computed from the raw
 *     store and the ICE state with the emission's per-entry arithmetic.  Full matrix on one device, after hb_ice_run. */
HB_API int hb_trans_order_stats(hb_csr* h, const int64_t* chr_offsets, int nchr, int64_t k, int64_t* n_trans_out,
                         double* v_k, double* v_k1);
HB_API int hb_trans_clip(hb_csr* h, const int64_t* chr_offsets, int nchr, double max_inter, int64_t* n_clipped_out);
HB_API int hb_ice_scaled_row_sums(hb_csr* h, double* out);

/* Step-level device API (all vectors are device fp64[n_cols], replicated per rank).
 *   hb_dev_ice_begin : bias = 1, r = 1, clear convergence state.
 *   hb_dev_ice_spmv  : exch[row0+i] = r_i * sum_j A_ij r_j for the block's rows and
 *                      exch[n_cols + rank] = max_ij A_ij r_i r_j over the block; every other
 *                      element of exch[0 .. n_cols+world) is zeroed first, so that ONE
 *                      sum-allreduce of exch over the ranks assembles the marginals.
 *   hb_dev_ice_update: from the assembled exch: per segment mean over non-zero
 *                      marginals, bias *= s, deviation, r /= s, convergence + overflow flags.
 *                      Identical on every rank (deterministic fixed-tile reductions).
 *   hb_dev_ice_status: async copy of {all_done, overflow, iteration} to a pinned int32[4].
 *   hb_dev_ice_finish: bias normalisation (lines 77-78); r is rescaled so that
 *                      A_ij r_i r_j is the final corrected value.
 *   hb_dev_ice_emit  : d_out[k] = A_k r_i r_j for the block (+ max into d_max[0], may be NULL). */
/*   hb_dev_ice_loop  : ONE cooperative launch that runs up to `iters` whole iterations (SpMV + update) back to back
 *                      on the device -- three grid barriers per iteration instead of kernel launches, no host in the
 *                      loop; stops early once every segment has converged, a value exceeds 1e100 or a peer times out.
 *                      Continues from the state hb_dev_ice_begin / earlier calls left; asynchronous.  Needs the
 *                      exchange on the device: a full matrix on one GPU, or row blocks in peer mode (hb_peer_attach).
 *                      Results are bit-identical to `iters` x (hb_dev_ice_spmv, hb_dev_ice_update).  hb_ice_run uses it
 *                      whenever it can (set HB_NO_PERSISTENT=1 in the environment to force the step-level path). */
HB_API int hb_dev_ice_loop(hb_csr* h, int iters, double tol, void* stream);
/* The 16 int32 progress words the persistent kernels mirror into pinned host memory while they run (no synchronisation):
 * [0] all done, [1] overflow, [2] iterations completed, [3] active segments, [4] exchanges of the last launch,
 * [5] KR outer iterations, [6] KR mat-vecs, [7] KR status, [8..15] diagnostics of a stage wait that gave up. */
HB_API int hb_loop_progress(hb_csr* h, int32_t* words16);
HB_API int hb_dev_ice_begin(hb_csr* h, void* stream);
HB_API int hb_dev_ice_spmv(hb_csr* h, double* d_exch, int rank, int world, void* stream);
HB_API int hb_dev_ice_update(hb_csr* h, const double* d_exch, int world, double tol, void* stream);
HB_API int hb_dev_ice_status(hb_csr* h, int32_t* pinned_status4, void* stream);
HB_API int hb_dev_ice_finish(hb_csr* h, void* stream);
HB_API int hb_dev_ice_emit(hb_csr* h, double* d_out, double* d_max, void* stream);
/* device pointers of the replicated state: bias[n_cols], r[n_cols] (1/bias, 0 for dead bins) */
HB_API int hb_dev_ice_vectors(hb_csr* h, double** d_bias, double** d_r);
/* host copy of per-segment results after the loop (sync) */
HB_API int hb_ice_results(hb_csr* h, double* bias_out, int32_t* iters_out, double* dev_out);

/* ---- Knight-Ruiz -------------------------------------------------------------
 * Replaces the krbalancing object (call sites hicCorrectMatrix.py:718-732, 744-757):
 *   kr_balancing(rows, cols, nnz, indptr, indices, data)  -> hb_csr_create(..., indices_i64=1)
 *   .computeKR()                                           -> hb_kr_run
 *   .get_normalisation_vector(rescale)                     -> x_out of hb_kr_run (+ hb_kr_rescale)
 *   .get_normalised_matrix(rescale)                        -> hb_kr_scaled_upper
 * The algorithm is Knight & Ruiz's `bnewt` on A + diag_eps*I with krbalancing's
 * constants (tol 1e-6, delta 0.1, Delta 3, g 0.9, etamax 0.1, diag_eps 1e-5).
 * x_out[n_cols]; outer_out / matvecs_out / residual_out optional. */
HB_API int hb_kr_run(hb_csr* h, double tol, double delta, double Delta, double diag_eps, int max_outer,
              double* x_out, int32_t* outer_out, int32_t* matvecs_out, double* residual_out);
/* factor = sqrt(sum_ij (A+eps I)_ij x_i x_j / sum_ij (A+eps I)_ij); x_inout /= factor (krbalancing rescale_norm_vector) */
HB_API int hb_kr_rescale(hb_csr* h, double diag_eps, double* x_inout, double* factor_out);
/* upper triangle (diagonal always present) of (A + eps I) o (x x^T):
 * call with all outputs NULL to get nnz_out, then with buffers indptr[n+1], indices[nnz], data[nnz]. */
HB_API int hb_kr_scaled_upper(hb_csr* h, double diag_eps, const double* x, int64_t* nnz_out, int64_t* indptr,
                       int32_t* indices, double* data);

/* step-level device API for the sharded KR loop: d_out[row0+i] = a_i * (sum_j A_ij u_j + eps*u_i) + b_i*c_i
 * for the block's rows (a, b, c optional device vectors indexed by GLOBAL row; NULL a -> 1, NULL b or c -> 0);
 * elements of d_out outside the block are zeroed when zero_rest != 0 (sum-allreduce assembly). */
HB_API int hb_dev_spmv(hb_csr* h, const double* d_u, double eps, const double* d_a, const double* d_b,
                const double* d_c, double* d_out, int zero_rest, void* stream);

/* ---- synthetic Hi-C matrices (benchmark inputs) --------------------
 * Deterministic symmetric matrix generated directly in HBM, block by block:
 * chromosome layout chrom_bins[nchrom]; cis entry (i,j) in the same chromosome with |i-j| <= band
 * exists iff hash(seed,min,max) < min(1, (d0/(|i-j|+1))^gamma); n_trans far entries per row (scale lambda_trans) at
 * j = (c_k - i) mod n; integer counts with planted per-bin biases, `low_frac` low-coverage
 * bins and `empty_frac` empty bins.  The same generator in numpy (tests/synth_ref.py) reproduces
 * it on the CPU for cross-checks.  Rows [row0, row0+n_rows). */
typedef struct hb_synth_params {
    uint64_t seed;
    int32_t nchrom;
    int32_t band;
    int32_t n_trans;
    int32_t reserved;
    double d0;
    double gamma;
    double lambda0;
    double low_frac;
    double empty_frac;
    double lambda_trans; /* mean scale of the far-entry counts */
} hb_synth_params;
HB_API int hb_synth_create(hb_csr** out, const hb_synth_params* p, const int64_t* chrom_bins, int64_t row0,
                    int64_t n_rows, int device);

#ifdef __cplusplus
}
#endif
#endif /* HICBALANCE_H */
