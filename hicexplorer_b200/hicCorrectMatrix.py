"""``hicCorrectMatrix correct`` on a B200: same sub-command, flags, defaults, log lines, exit codes and output
files as the reference driver (hicexplorer/hicCorrectMatrix.py:32-254, 595-779), with the numerics on the GPU.

Differences, all additive or documented:
  * ``--tolerance`` (default 1e-5 = the value hard-wired in the reference, iterativeCorrection.py:10) and
    ``--device`` are new optional flags;
  * the ICE branch uploads the matrix once and runs zero-bin masking, NaN/Inf scrub, the MAD filter, bin deletion
    and the iterations on the device (``pipeline.ice_pipeline``) instead of round-tripping through scipy;
  * launched under ``torchrun`` (WORLD_SIZE > 1) the genome-wide ICE / KR corrections run on nnz-balanced row blocks,
    one process per GPU, ``--perchr`` KR deals whole chromosomes to the GPUs, and rank 0 writes the output
    (``--sequencedCountCutoff``, ``--transCutoff`` and ``--inflationCutoff`` work on the row blocks too);
  * ``diagnostic_plot`` (:425-545) computes its statistics on the device -- coverage per bin, modified z-scores, the
    100-bin histogram of the bins with z < 5 and the z-score of its first local minimum, logged as "mad threshold" like
    the reference -- and draws the histogram with matplotlib when it is installed, else as a hand-written SVG;
  * ``--transCutoff`` computes hicmatrix ``truncTrans``'s percentile of the trans values on the device and keeps the
    row sums for ``--inflationCutoff``; like the published hicmatrix code (whose clipping statement is a comparison)
    it leaves the values unchanged unless the additive ``--clipTrans`` flag asks for the documented clipping.
    ``--inflationCutoff`` only ever worked together with ``--transCutoff`` in the reference (``pre_row_sum`` is
    otherwise undefined, :699/:771 -> NameError) and exits with an explanation when given alone.
"""
import argparse
import logging
import sys

import numpy as np
from scipy import sparse
from scipy.sparse import lil_matrix

from . import hicmatrix_lite as hm
from ._version import __version__
from .krbalancing import kr_balancing
from .pipeline import ice_pipeline, ice_pipeline_sharded, kr_pipeline, kr_pipeline_sharded

log = logging.getLogger(__name__)


def parse_arguments(args=None):
    parser = argparse.ArgumentParser(
        formatter_class=argparse.RawDescriptionHelpFormatter, conflict_handler='resolve',
        description="""
This function provides 2 balancing methods which can be applied on a raw matrix.

I. KR: It balances a matrix using a fast balancing algorithm introduced by Knight and Ruiz (2012).

II. ICE: Iterative correction of a Hi-C matrix (see Imakaev et al. 2012). For this method to work correctly, bins
with zero reads assigned to them should be removed as they cannot be corrected, as well as bins with low or
extremely high coverage (--filterThreshold).

B200-native implementation: the correction runs on the GPU through libhicbalance.so.
""")
    parser.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    subparsers = parser.add_subparsers(title="Options", dest='command', metavar='', help="""To get detailed help on each of the options: \r
    $ hicCorrectMatrix diagnostic_plot -h \r
    $ hicCorrectMatrix correct -h """)
    plot_mode = subparsers.add_parser('diagnostic_plot', formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                                      help="Plots a histogram of the coverage per bin together with the modified "
                                           "z-score based on the median absolute deviation method.",
                                      usage='%(prog)s --matrix hic_matrix.h5 -o file.png')
    plot_mode.add_argument('--matrix', '-m', help='Name of the Hi-C matrix to correct in .h5 format.', required=True)
    plot_mode.add_argument('--plotName', '-o', help='File name to save the diagnostic plot.', required=True)
    plot_mode.add_argument('--chromosomes', help='List of chromosomes to be included in the iterative correction.',
                           default=None, nargs='+')
    plot_mode.add_argument('--xMax', help='Max value for the x-axis in counts per bin.', default=None, type=float)
    plot_mode.add_argument('--perchr', help='Compute histogram per chromosome.', action='store_true')
    plot_mode.add_argument('--verbose', help='Print processing status.', action='store_true')
    plot_mode.add_argument('--device', help='CUDA device index.', type=int, default=0)
    subparsers.add_parser('correct', formatter_class=argparse.ArgumentDefaultsHelpFormatter, parents=[correct_subparser()],
                          help="""Run Knight-Ruiz matrix balancing algorithm (KR) or the iterative matrix correction (ICE).""",
                          usage='%(prog)s --matrix hic_matrix.h5 --filterThreshold -1.2 5 (Only if ICE) -out corrected_matrix.h5 \n')
    return parser


def correct_subparser():
    parser = argparse.ArgumentParser(add_help=False)
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrix', '-m', help='Name of the Hi-C matrix to correct in .h5 format.', required=True)
    req.add_argument('--outFileName', '-o', help='File name to save the resulting matrix. The output is a .h5 file.',
                     required=True)
    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--correctionMethod', help='Method to be used for matrix correction. It can be set to KR or ICE'
                     ' (Default: %(default)s).', type=str, metavar='STR', default='KR', choices=['KR', 'ICE'])
    opt.add_argument('--filterThreshold', '-t', help='Removes bins of low or large coverage. A lower and upper '
                     'threshold are required separated by space, e.g. --filterThreshold -1.5 5. Applied only for ICE!',
                     type=float, nargs=2, default=None)
    opt.add_argument('--iterNum', '-n', help='Number of iterations to compute. only for ICE! (Default: %(default)s).',
                     type=int, metavar='INT', default=500)
    opt.add_argument('--inflationCutoff', help='Value corresponding to the maximum number of times a bin can be '
                     'scaled up during the iterative correction. Only for ICE!', type=float)
    opt.add_argument('--transCutoff', '-transcut', help='Clip high counts in the top -transcut trans regions '
                     '(i.e. between chromosomes). Only for ICE!', type=float)
    opt.add_argument('--sequencedCountCutoff', help='Each bin receives a value indicating the fraction that is '
                     'covered by reads. A cutoff of 0.5 will discard all those bins that have less than half of the '
                     'bin covered. Only for ICE!', default=None, type=float)
    opt.add_argument('--chromosomes', help='List of chromosomes to be included in the iterative correction. The '
                     'order of the given chromosomes will be then kept for the resulting corrected matrix',
                     default=None, nargs='+')
    opt.add_argument('--skipDiagonal', '-s', help='If set, diagonal counts are not included. Only for ICE!',
                     action='store_true')
    opt.add_argument('--perchr', help='Normalize each chromosome separately. This is useful for samples from cells '
                     'with uneven number of chromosomes and/or translocations.', action='store_true')
    opt.add_argument('--filteredBed', help='Print bins filtered our by  --filterThreshold to this file')
    opt.add_argument('--verbose', help='Print processing status.', action='store_true')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    # additive flags of the B200 build
    opt.add_argument('--tolerance', help='ICE convergence tolerance on the marginals (Default: %(default)s, the '
                     'value hard-wired in the reference).', type=float, default=1e-5)
    opt.add_argument('--device', help='CUDA device index (Default: %(default)s).', type=int, default=0)
    opt.add_argument('--clipTrans', help='With --transCutoff: clip the trans values at or above the percentile to it '
                     '(what truncTrans documents; the published hicmatrix statement is a no-op comparison).',
                     action='store_true')
    return parser


def _is_cooler(path):
    fname = str(path).partition('::')[0]
    return fname.endswith('.cool') or fname.endswith('.mcool')


def main(args=None):
    args = parse_arguments().parse_args(args)
    if args.verbose:
        log.setLevel(logging.INFO)
        logging.getLogger('hicexplorer_b200.pipeline').setLevel(logging.INFO)
    if getattr(args, 'command', None) == 'diagnostic_plot' or 'plotName' in args:
        if _is_cooler(args.matrix) and args.chromosomes is not None and len(args.chromosomes) == 1:
            ma = hm.hiCMatrix(args.matrix, pChrnameList=[str(c) for c in args.chromosomes])
        else:
            ma = hm.hiCMatrix(args.matrix)
            if args.chromosomes:
                ma.reorderChromosomes([str(c) for c in args.chromosomes])
        log.setLevel(logging.INFO)
        panels = diagnostic_statistics(ma, perchr=args.perchr, x_max=args.xMax, device=args.device)
        draw_diagnostic_plot(panels, args.plotName)
        log.info("Saving diagnostic plot {}\n".format(args.plotName))
        return panels

    world, rank, own_group = _init_distributed(args)
    if _is_cooler(args.matrix) and args.chromosomes is not None and len(args.chromosomes) == 1:
        ma = hm.hiCMatrix(args.matrix, pChrnameList=[str(c) for c in args.chromosomes])
    else:
        ma = hm.hiCMatrix(args.matrix)
        if args.chromosomes:
            ma.reorderChromosomes([str(c) for c in args.chromosomes])

    if args.correctionMethod == 'ICE':
        _correct_ice(ma, args, world, rank)
    else:
        _correct_kr(ma, args, world, rank)
    if rank == 0:
        ma.save(args.outFileName, pApplyCorrection=False)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        if own_group:
            dist.destroy_process_group()


def diagnostic_statistics(ma, perchr=False, x_max=None, device=0):
    """What ``diagnostic_plot`` shows (hicCorrectMatrix.py:425-545), as numbers.  Coverage (row sum minus diagonal, per
    chromosome block with ``perchr``), medians / MADs and modified z-scores come from the device (hb_row_stats[_cis],
    hb_mad_filter); the 100-bin histogram of the bins with z < 5 and its first local minimum are O(n) host work.
    Returns one dict per histogram: title, values, counts, edges, threshold, mad_threshold, median, mad."""
    from .store import DeviceCSR
    up = ma.upper_for_device()
    with (DeviceCSR.from_upper(up, device=device) if up is not None else DeviceCSR.from_scipy(ma.matrix, device=device)) as dev:
        drop = ma.upper_drop_mask() if up is not None else None
        if drop is not None:
            dev.compact(drop)
        # "mask all zero value bins" happens for the plot too (:616-619), on row sums that include the diagonal and
        # still see NaNs; the medians, z-scores and the histogram are those of the remaining bins
        row_sum_all, _ = dev.row_stats()
        zero = row_sum_all == 0
        log.info("Removing {} zero value bins".format(int(zero.sum())))
        chrom_offsets = ma.chromosome_offsets()
        if zero.any():
            dev.compact(zero.astype(np.uint8))
            from .pipeline import _segment_offsets_after
            chrom_offsets = _segment_offsets_after(zero, chrom_offsets)
        dev.scrub()                                         # convertNansToZeros / convertInfsToZeros (:624-625, :486-487)
        offsets = np.asarray(chrom_offsets, dtype=np.int64) if perchr else np.asarray([0, dev.n], dtype=np.int64)
        if perchr:
            dev.set_segments(offsets)
        row_sum, diag = dev.row_stats(cis_only=perchr)
        _mask, z, med, mad = dev.mad_filter(-np.inf, np.inf, want_z=True)
    points = row_sum - diag
    names = ma.getChrNames() if perchr else [None]
    if perchr and len(names) > 30:
        log.warning("The matrix contains {} chromosomes. It is not practical to plot each. Try using --chromosomes to "
                    "select some chromosomes or plot a single histogram.".format(len(names)))
    panels = []
    for k, title in enumerate(names):
        lo, hi = int(offsets[k]), int(offsets[k + 1])
        if title is not None:
            log.info("Plotting chromosome {}".format(title))
        values = points[lo:hi][z[lo:hi] < 5]                # high remove outliers (:510, :529)
        panel = dict(title=title, values=values, counts=None, edges=None, threshold=None, mad_threshold=None,
                     median=float(med[k]), mad=float(mad[k]))
        if len(values) == 0:
            if title is None:
                log.error('No data available, exit.')
                sys.exit(1)
            log.error('No data for chromosome {}. Continue with next chromosome.'.format(title))
        if x_max:
            values = values[values < x_max]
        if len(values):
            counts, edges = np.histogram(values, 100)       # ax1.hist(row_sum_values, 100)
            local_min = [x for x, y in enumerate(counts) if 1 <= x < len(counts) - 1 and counts[x - 1] > y < counts[x + 1]]
            threshold = edges[local_min[0]] if local_min else None
            panel.update(values=values, counts=counts, edges=edges, threshold=threshold)
            if threshold:
                diff = threshold - panel["median"]
                mad_threshold = 0.6745 * diff if panel["mad"] == 0.0 else 0.6745 * diff / panel["mad"]
                panel["mad_threshold"] = mad_threshold
                if title:
                    log.info("{}: mad threshold {}".format(title, mad_threshold))
                else:
                    log.info("mad threshold {}".format(mad_threshold))
        panels.append(panel)
    return panels


def _value_to_mad(panel, value):
    diff = np.asarray(value, dtype=np.float64) - panel["median"]
    return 0.6745 * diff if panel["mad"] == 0.0 else 0.6745 * diff / panel["mad"]


def draw_diagnostic_plot(panels, path):
    """histogram of the coverage per bin with the modified z-score on the top axis and the first local minimum marked;
    matplotlib if it is installed (same figure layout as the reference), otherwise a plain SVG written by hand"""
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except ImportError:
        return _draw_svg(panels, path)
    cols = min(len(panels), 5)
    rows = int(np.ceil(len(panels) / 5.0))
    fig = plt.figure(figsize=(6 * cols, 5 * rows)) if len(panels) > 1 else plt.figure()
    for k, panel in enumerate(panels):
        ax1 = fig.add_subplot(rows, cols, k + 1)
        ax1.set_xlabel("total counts per bin")
        ax1.set_ylabel("frequency")
        if panel["counts"] is not None:
            ax1.bar(panel["edges"][:-1], panel["counts"], width=np.diff(panel["edges"]), align='edge', color='green')
            ax2 = ax1.twiny()
            ax2.set_xlabel("modified z-score")
            ax2.set_xlim(_value_to_mad(panel, np.array(ax1.get_xlim())))
            if panel["mad_threshold"] is not None:
                ymin, ymax = ax2.get_ylim()
                ax2.vlines(panel["mad_threshold"], ymin, ymax)
        if panel["title"]:
            ax1.set_title(panel["title"])
    plt.tight_layout()
    plt.savefig(path)
    plt.close()


def _draw_svg(panels, path):
    if not str(path).lower().endswith('.svg'):
        log.warning("matplotlib is not installed: writing the plot as SVG to {}.svg".format(path))
        path = str(path) + '.svg'
    cols = min(len(panels), 5)
    rows = int(np.ceil(len(panels) / 5.0))
    pw, ph, ml, mt, mb = 420, 320, 55, 45, 40
    out = ['<svg xmlns="http://www.w3.org/2000/svg" width="{}" height="{}" font-family="sans-serif" font-size="11">'.format(
        cols * pw, rows * ph), '<rect width="100%" height="100%" fill="white"/>']
    for k, panel in enumerate(panels):
        x0, y0 = (k % 5) * pw + ml, (k // 5) * ph + mt
        w, h = pw - ml - 15, ph - mt - mb
        out.append('<rect x="{}" y="{}" width="{}" height="{}" fill="none" stroke="black"/>'.format(x0, y0, w, h))
        if panel["title"]:
            out.append('<text x="{}" y="{}" text-anchor="middle" font-weight="bold">{}</text>'.format(x0 + w / 2, y0 - 30, panel["title"]))
        out.append('<text x="{}" y="{}" text-anchor="middle">total counts per bin</text>'.format(x0 + w / 2, y0 + h + 32))
        out.append('<text x="{}" y="{}" text-anchor="middle">modified z-score</text>'.format(x0 + w / 2, y0 - 18))
        if panel["counts"] is None:
            continue
        edges, counts = panel["edges"], panel["counts"]
        lo, hi, top = float(edges[0]), float(edges[-1]), float(max(counts.max(), 1))
        span = (hi - lo) or 1.0
        for c, a, b in zip(counts, edges[:-1], edges[1:]):
            bh = h * float(c) / top
            out.append('<rect x="{:.2f}" y="{:.2f}" width="{:.2f}" height="{:.2f}" fill="green"/>'.format(
                x0 + w * (a - lo) / span, y0 + h - bh, max(w * (b - a) / span, 0.5), bh))
        for frac in (0.0, 0.25, 0.5, 0.75, 1.0):
            v = lo + frac * span
            out.append('<text x="{:.1f}" y="{}" text-anchor="middle">{:.4g}</text>'.format(x0 + w * frac, y0 + h + 14, v))
            out.append('<text x="{:.1f}" y="{}" text-anchor="middle">{:.2f}</text>'.format(x0 + w * frac, y0 - 5, float(_value_to_mad(panel, v))))
        out.append('<text x="{}" y="{}" text-anchor="end">{:.0f}</text>'.format(x0 - 4, y0 + 10, top))
        if panel["threshold"] is not None and panel["mad_threshold"] is not None:
            xt = x0 + w * (float(panel["threshold"]) - lo) / span
            out.append('<line x1="{0:.2f}" y1="{1}" x2="{0:.2f}" y2="{2}" stroke="black"/>'.format(xt, y0, y0 + h))
            out.append('<text x="{:.1f}" y="{}" font-size="10">z = {:.2f}</text>'.format(xt + 3, y0 + 12, panel["mad_threshold"]))
    out.append('</svg>')
    with open(path, 'w') as fh:
        fh.write("\n".join(out))


def _init_distributed(args):
    """one process per GPU when launched by torchrun; (world, rank, group created here) = (1, 0, False) otherwise"""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 1, 0, False
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    args.device = local
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, dist.get_rank(), created


def _correct_ice(ma, args, world=1, rank=0):
    if not args.filterThreshold:
        log.error('min and max filtering thresholds should be set')
        sys.exit(1)
    trans_on = bool(args.transCutoff and 0 < args.transCutoff < 100)
    if args.inflationCutoff and args.inflationCutoff > 0 and not trans_on:
        log.error('--inflationCutoff needs --transCutoff (the reference only defines pre_row_sum then, '
                  'hicCorrectMatrix.py:699, 771)')
        sys.exit(1)
    intervals_all = list(ma.cut_intervals)
    extra_fn = None
    if args.sequencedCountCutoff and 0 < args.sequencedCountCutoff < 1:
        coverage = np.array([iv[3] for iv in intervals_all], dtype=np.float64)

        def extra_fn(kept_ids):
            return coverage[kept_ids] < args.sequencedCountCutoff

    if world > 1:
        file_layout = not args.perchr or str(args.outFileName).endswith('.h5')
        # every rank maps the same file; when the matrix is untouched each uploads the stored upper triangle, builds the
        # symmetric store on its GPU and keeps its own rows -- no symmetric matrix on any host
        up = ma.upper_for_device()
        res = ice_pipeline_sharded(up if up is not None else ma.matrix, (args.filterThreshold[0], args.filterThreshold[1]),
                                   max_iter=args.iterNum, tolerance=args.tolerance, chr_offsets=ma.chromosome_offsets(),
                                   perchr=args.perchr, skip_diagonal=args.skipDiagonal, want_matrix=file_layout,
                                   file_layout=file_layout, upper_input=up is not None,
                                   drop_mask=ma.upper_drop_mask() if up is not None else None, extra_mask_fn=extra_fn,
                                   trans_cutoff=args.transCutoff, clip_trans=args.clipTrans,
                                   inflation_cutoff=args.inflationCutoff)
        if rank != 0:
            return
        if res.get('max_inter') is not None:
            log.info("trans contacts: {} percentile = {}".format(100 - float(args.transCutoff) / 100, res['max_inter']))
        if 'inflated_rel' in res:
            log.info("inflated >={} regions: {}".format(args.inflationCutoff, len(res['inflated_rel'])))
    else:
        # the file's upper triangle goes to the device as it is (the mirror image is built in HBM) and the corrected
        # matrix comes back already in file layout, whenever nothing on the host has touched the matrix since loading
        up = ma.upper_for_device()
        file_layout = not args.perchr or str(args.outFileName).endswith('.h5')
        res = ice_pipeline(up if up is not None else ma.matrix, (args.filterThreshold[0], args.filterThreshold[1]),
                           max_iter=args.iterNum, tolerance=args.tolerance, chr_offsets=ma.chromosome_offsets(),
                           perchr=args.perchr, skip_diagonal=args.skipDiagonal, device=args.device, want_matrix=file_layout,
                           extra_mask_fn=extra_fn, upper_input=up is not None, file_layout=file_layout,
                           trans_cutoff=args.transCutoff, clip_trans=args.clipTrans, inflation_cutoff=args.inflationCutoff,
                           drop_mask=ma.upper_drop_mask() if up is not None else None)
        if res.get('max_inter') is not None:
            log.info("trans contacts: {} percentile = {}".format(100 - float(args.transCutoff) / 100, res['max_inter']))
        if 'inflated_rel' in res:
            log.info("inflated >={} regions: {}".format(args.inflationCutoff, len(res['inflated_rel'])))
    n = res['n']
    keep1 = np.setdiff1d(np.arange(n), res['zero_bins'])
    outlier_regions = res['outliers_rel']
    pct_outlier = 100 * float(len(outlier_regions)) / max(len(keep1), 1)
    log.info("Bins that are MAD outliers ({:.2f}%) out of {}: {}".format(pct_outlier, len(keep1), len(outlier_regions)))
    if args.filteredBed:
        with open(args.filteredBed, 'w') as f:
            # the reference iterates a Python set of numpy integers here (hicCorrectMatrix.py:668)
            for outlier_region in set(np.int64(v) for v in outlier_regions):
                interval = intervals_all[keep1[outlier_region]]
                f.write(f"{interval[0]}\t{interval[1]}\t{interval[2]}\t.\t{interval[3]}\t.\n")
    iters = res['iterations']
    for k, it in enumerate(iters):
        if res['deviation'][k] < args.tolerance:
            log.info("[iterative correction] {} iterations used\n".format(int(it) + 1))
    if 'upper' in res:
        # corrected upper triangle in the original frame + NaN-padded factors: what setMatrixValues,
        # setCorrectionFactors and restoreMaskedBins (inside save) would assemble on the host
        factors = np.full(n, np.nan)
        factors[res['kept']] = res['bias']
        ma.install_file_layout(res['upper'], np.setdiff1d(np.arange(n), res['kept']), factors)
    else:
        # install the corrected matrix / factors in the kept frame; save() re-inserts the masked bins as nan_bins
        if 'matrix' in res:
            ma.apply_mask_bookkeeping(res['kept'], matrix=res['matrix'])
        else:
            ma.apply_mask_bookkeeping(res['kept'])     # perchr + .cool keeps the raw counts (hicCorrectMatrix.py:759-760)
        ma.setCorrectionFactors(res['bias'])
    log.info("Total regions to be removed: {}".format(n - len(res['kept'])))


def _block_diag_upper(blocks, offsets, n):
    """file-layout CSR (n x n) from per-chromosome upper-triangular blocks placed on the diagonal"""
    indptr = np.zeros(n + 1, dtype=np.int64)
    idx_parts, val_parts = [], []
    base = 0
    for (lo, hi), blk in zip(zip(offsets[:-1], offsets[1:]), blocks):
        blk = sparse.csr_matrix(blk)
        indptr[lo + 1:hi + 1] = base + blk.indptr[1:]
        idx_parts.append(blk.indices.astype(np.int32, copy=False) + np.int32(lo))
        val_parts.append(blk.data)
        base += blk.nnz
    indptr[offsets[-1] + 1:] = base
    idx = np.concatenate(idx_parts) if idx_parts else np.zeros(0, dtype=np.int32)
    val = np.concatenate(val_parts) if val_parts else np.zeros(0)
    return sparse.csr_matrix((val, idx, indptr), shape=(n, n))


def _drop_explicit_zeros(m):
    if m.nnz and np.any(m.data == 0):
        m.eliminate_zeros()
    return m


def _correct_kr_from_upper(ma, up, args):
    """KR when the loaded matrix is untouched (single GPU): the file's upper triangle is uploaded as it is, the symmetric
    store is built on the device and the balanced matrix comes back in file layout; the full symmetric matrix never
    exists on the host.  Same results as the kr_balancing object protocol of the reference (:718-732, :744-757)."""
    want_matrix = str(args.outFileName).endswith('.h5')
    drop = ma.upper_drop_mask()
    if drop is not None and (args.perchr or not want_matrix):
        # the raw counts of the selected chromosomes are needed on the host (blocks / .cool pixels): select them there
        keep = np.flatnonzero(drop == 0)
        up = sparse.csr_matrix(up[keep, :][:, keep])
        up.sort_indices()
        drop = None
    n = len(ma.cut_intervals)
    if not args.perchr:
        res = kr_pipeline(up, rescale=True, want_matrix=want_matrix, device=args.device, upper_input=True, drop_mask=drop,
                          skip_diagonal=args.skipDiagonal)
        print("normalisation factor is {}".format(res['factor']))
        upper = _drop_explicit_zeros(res['upper']) if want_matrix else up
        ma.install_file_layout(upper, ma.nan_bins, np.asarray(res['x']).reshape(-1, 1))
        return
    offsets = ma.chromosome_offsets()

    def balance(ci):
        lo, hi = int(offsets[ci]), int(offsets[ci + 1])
        res = kr_pipeline(up[lo:hi, lo:hi].tocsr(), rescale=want_matrix, want_matrix=want_matrix, device=args.device,
                          upper_input=True, skip_diagonal=args.skipDiagonal)
        if want_matrix:
            print("normalisation factor is {}".format(res['factor']))
        return ci, res

    from concurrent.futures import ThreadPoolExecutor
    order = sorted(range(len(offsets) - 1), key=lambda ci: -(offsets[ci + 1] - offsets[ci]))
    with ThreadPoolExecutor(max_workers=4) as pool:
        done = dict(pool.map(balance, order))
    factors = np.concatenate([np.asarray(done[ci]['x']).reshape(-1, 1) for ci in range(len(offsets) - 1)])
    if want_matrix:
        upper = _drop_explicit_zeros(_block_diag_upper([done[ci]['upper'] for ci in range(len(offsets) - 1)], offsets, n))
    else:
        upper = up
    ma.install_file_layout(upper, ma.nan_bins, factors)


def _scrubbed(matrix, skip_diagonal):
    """what both methods see (hicCorrectMatrix.py:624-626, 648-649): NaN / Inf -> 0, float64, optionally a zero diagonal
    (host-side route of re-ordered matrices; the device routes call hb_scrub / hb_diag_zero)"""
    m = sparse.csr_matrix(matrix, dtype=np.float64, copy=True)
    m.data[~np.isfinite(m.data)] = 0
    if skip_diagonal:
        m.setdiag(0)
    return m


def _correct_kr(ma, args, world=1, rank=0):
    up = ma.upper_for_device()
    if up is not None and world == 1:
        return _correct_kr_from_upper(ma, up, args)
    if world > 1 and not args.perchr:
        # genome-wide KR on nnz-balanced row blocks over the GPUs: the (rescaled) vector is replicated; for .h5 every rank
        # also emits its rows of the balanced matrix in file layout and rank 0 concatenates them
        want_matrix = str(args.outFileName).endswith('.h5')
        if up is not None:
            drop = ma.upper_drop_mask()
            res = kr_pipeline_sharded(up, rescale=True, upper_input=True, drop_mask=drop, skip_diagonal=args.skipDiagonal,
                                      want_matrix=want_matrix)
            if rank == 0:
                print("normalisation factor is {}".format(res['factor']))
                if want_matrix:
                    out_upper = _drop_explicit_zeros(res['upper'])
                else:
                    out_upper = up
                    if drop is not None:              # the .cool keeps the raw counts of the selected chromosomes
                        keep = np.flatnonzero(drop == 0)
                        out_upper = sparse.csr_matrix(up[keep, :][:, keep])
                        out_upper.sort_indices()
                ma.install_file_layout(out_upper, ma.nan_bins, np.asarray(res['x']).reshape(-1, 1))
            return
        res = kr_pipeline_sharded(_scrubbed(ma.matrix, args.skipDiagonal), rescale=True, want_matrix=want_matrix)
        if rank == 0:
            print("normalisation factor is {}".format(res['factor']))
            if want_matrix:
                full = res['upper'] + sparse.triu(res['upper'], 1).T
                ma.setMatrixValues(sparse.csr_matrix(full))
            ma.setCorrectionFactors(np.asarray(res['x']).reshape(-1, 1))
        return
    matrix = _scrubbed(ma.matrix, args.skipDiagonal)
    if args.perchr:
        # independent chromosomes: dealt to the GPUs by longest-processing-time on stored entries, no communication
        # during the correction; rank 0 collects the factors (and the corrected blocks for .h5)
        from .dist import lpt_assignment
        want_matrix = str(args.outFileName).endswith('.h5')
        names = ma.getChrNames()
        ranges = [ma.getChrBinRange(c) for c in names]
        owner = lpt_assignment([int(matrix.indptr[hi] - matrix.indptr[lo]) for lo, hi in ranges], world)
        def balance(ci):
            lo, hi = ranges[ci]
            sub = matrix[lo:hi, lo:hi].tocsr()
            kr = kr_balancing(sub.shape[0], sub.shape[1], sub.count_nonzero(), sub.indptr.astype(np.int64, copy=False),
                              sub.indices.astype(np.int64, copy=False), sub.data.astype(np.float64, copy=False),
                              device=args.device)
            kr.computeKR()
            block = kr.get_normalised_matrix(True) if want_matrix else None
            vec = kr.get_normalisation_vector(False).todense()
            kr.close()
            return ci, (block, vec)

        # a few host threads, each running whole chromosomes on its own handle / stream (the C ABI releases the GIL):
        # the per-mat-vec host synchronisation of one chromosome hides behind the kernels of another
        # (tools/perchr_kr_threads.py: 4.8 -> 6.4 TB/s on the 5 kb matrix); results do not depend on the thread count
        from concurrent.futures import ThreadPoolExecutor
        todo = sorted((ci for ci in range(len(ranges)) if owner[ci] == rank),
                      key=lambda ci: -(ranges[ci][1] - ranges[ci][0]))
        with ThreadPoolExecutor(max_workers=4) as pool:
            mine = dict(pool.map(balance, todo))
        if world > 1:
            import torch.distributed as dist
            parts = [None] * world if rank == 0 else None
            dist.gather_object(mine, parts, dst=0)
            if rank != 0:
                return
            mine = {}
            for p in parts:
                mine.update(p)
        corrected_matrix = lil_matrix(matrix.shape)
        factors = []
        for ci, (lo, hi) in enumerate(ranges):
            block, vec = mine[ci]
            if want_matrix:
                corrected_matrix[lo:hi, lo:hi] = block
            factors.append(vec)
        correction_factors = np.concatenate(factors)
    else:
        kr = kr_balancing(matrix.shape[0], matrix.shape[1], matrix.count_nonzero(),
                          matrix.indptr.astype(np.int64, copy=False), matrix.indices.astype(np.int64, copy=False),
                          matrix.data.astype(np.float64, copy=False), device=args.device)
        kr.computeKR()
        correction_factors = kr.get_normalisation_vector(True).todense()
        corrected_matrix = kr.get_normalised_matrix(True) if str(args.outFileName).endswith('.h5') else None
        kr.close()
    if str(args.outFileName).endswith('.h5'):
        ma.setMatrixValues(corrected_matrix)
    ma.setCorrectionFactors(correction_factors)


if __name__ == "__main__":
    main()
