#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that are not bench.py's headline line:

  perchr5k   configs[3]: --perchr ICE + KR on a synthetic 5 kb human matrix: 24 independent chromosome blocks,
             ICE batched as segments of ONE device matrix per rank (one launch covers all its chromosomes),
             KR chromosome by chromosome; chromosomes are dealt to the ranks by LPT on stored entries,
             no inter-GPU communication.
  kr1k       configs[4]: KR on a synthetic genome-wide 1 kb matrix (~3.1 M bins, ~8e9 entries = 96 GB of CSR),
             nnz-balanced row blocks over the ranks, one sum-allreduce of n fp64 per mat-vec.  Needs 8 GPUs.
  kr25k_sharded / ice10k_conv   the single-GPU configs run sharded (for N = 2, 4 checks).

Run: python tools/bench_configs.py CONFIG            (1 GPU)
     python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py CONFIG
Prints one JSON line per config on rank 0."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.WORKLOADS.update({
    "perchr5k": ("ICE+KR", 5000, 2.0e9, 16000),
    "kr1k": ("KR", 1000, 8.0e9, 8000),
})


def lpt(weights, world):
    """longest-processing-time assignment of chromosomes to ranks"""
    loads = [0.0] * world
    owner = [0] * len(weights)
    for c in sorted(range(len(weights)), key=lambda i: -weights[i]):
        r = int(np.argmin(loads))
        owner[c] = r
        loads[r] += weights[c]
    return owner


def main():
    import torch
    import torch.distributed as dist
    config = sys.argv[1] if len(sys.argv) > 1 else "perchr5k"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(config)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run(config):
    """one configuration on the already initialised process group (or a single GPU); every rank calls it, the returned
    dict is complete on rank 0"""
    import torch
    import torch.distributed as dist
    from hicexplorer_b200 import _lib
    from hicexplorer_b200.dist import attach_comm, attach_peer, nnz_balanced_partition
    from hicexplorer_b200.store import DeviceCSR
    L = _lib.lib()
    check = _lib.check
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = bench.measured_peak()
    out = {"config": config, "n_gpus": world, "peak_GBps": peak, "peak_source": peak_src}

    if config == "merge10k":
        # hicMergeMatrixBins --numBins 5 on the synthetic 10 kb matrix (10 kb -> 50 kb), 1 GPU
        method, bins, kw = bench.synth_kwargs("ice10k", clean=True)
        n = int(sum(bins))
        iv_chrom = np.repeat(np.arange(len(bins)), bins)
        k = 5
        # groups of k inside each chromosome (a trailing group < k/2 bins is dropped, like the reference)
        bin_map = np.full(n, -1, dtype=np.int32)
        nxt = 0
        start = 0
        for c, L in enumerate(bins):
            full, rest = divmod(L, k)
            ids = np.repeat(np.arange(full), k) + nxt
            bin_map[start:start + full * k] = ids
            nxt += full
            if rest and not rest < k / 2:
                bin_map[start + full * k:start + L] = nxt
                nxt += 1
            start += L
        if bin_map[-1] < 0:                       # the very last group of the matrix is always kept
            last = np.flatnonzero(bin_map >= 0)[-1] + 1
            bin_map[last:] = nxt
            nxt += 1
        times = []
        for rep in range(3):
            with DeviceCSR.synthetic(bins, device=local, **kw) as dev:
                nnz_in = dev.nnz
                dev.sync()
                t0 = time.time()
                dev.merge_bins(bin_map, nxt)
                dev.sync()
                times.append(time.time() - t0)
                nnz_out, n_out = dev.nnz, dev.n
        t = min(times)
        from oracle import hic_cpu, merge_oracle
        sample = hic_cpu.synth_matrix([bins[0]], **kw)
        siv = [("chr1", i, i + 1, 1.0) for i in range(bins[0])]
        t0 = time.time()
        ref_m, _iv, _nan, _map = merge_oracle.merge_bins(sample, siv, k)
        t_cpu = time.time() - t0
        out.update({"workload": "merge 5 bins (10 kb -> 50 kb) of the synthetic 10 kb matrix: %d bins, %d entries -> %d bins, %d entries"
                                % (n, nnz_in, n_out, nnz_out),
                    "merge": {"ms": t * 1e3, "input_entries_per_s": nnz_in / t,
                              "GBps": (2 * 12.0 * nnz_in + 12.0 * nnz_out) / t / 1e9,
                              "frac_of_peak": (2 * 12.0 * nnz_in + 12.0 * nnz_out) / t / 1e9 / peak,
                              "note": "two passes over the input (count, fill) + allocation of the merged store"},
                    "cpu_baseline": {"value": sample.nnz / t_cpu, "unit": "input entries/s", "cores": 1, "kind": "port",
                                     "sample": "oracle/merge_oracle.py (numpy/scipy restatement of reduce_matrix) on chromosome 1 "
                                               "of the same matrix (%d entries), %.1f s" % (sample.nnz, t_cpu)}})
    elif config == "elem10k":
        # hicNormalize (norm_range) and hicSumMatrices (A + B) on the synthetic 10 kb matrix, 1 GPU
        method, bins, kw = bench.synth_kwargs("ice10k", clean=True)
        from oracle import hic_cpu, normalize_oracle

        def wall(fn, dev):
            dev.sync()
            t0 = time.time()
            r = fn()
            dev.sync()
            return r, time.time() - t0
        res = {}
        with DeviceCSR.synthetic(bins, device=local, **kw) as dev:
            nnz = dev.nnz
            dev.value_stats()                                   # warm-up (lazy module load)
            (tot, mn, mx), t_stats = wall(lambda: dev.value_stats(scrub=True), dev)
            _, t_norm = wall(lambda: dev.normalize("norm_range", a=mn, b=np.float32(mx) - np.float32(mn)), dev)
            _, t_elim = wall(dev.eliminate_zeros, dev)
            nnz_after = dev.nnz
            res["normalize"] = {"value_stats_ms": t_stats * 1e3, "normalize_ms": t_norm * 1e3, "eliminate_zeros_ms": t_elim * 1e3,
                                "entries_in": nnz, "entries_out": nnz_after,
                                "stats_GBps": 8.0 * nnz / t_stats / 1e9, "normalize_GBps": 16.0 * nnz / t_norm / 1e9,
                                "eliminate_GBps": (8.0 * nnz + 12.0 * nnz + 12.0 * nnz_after) / t_elim / 1e9,
                                "entries_per_s_whole_tool": nnz / (t_stats + t_norm + t_elim),
                                "frac_of_peak_normalize": 16.0 * nnz / t_norm / 1e9 / peak}
        # hicTransform --method obs_exp --perChromosome: per-distance sums, then every entry over its expected value
        from hicexplorer_b200.hicTransform import expected_values
        with DeviceCSR.synthetic(bins, device=local, **kw) as dev:
            offs = np.concatenate([[0], np.cumsum(bins)]).astype(np.int64)
            dev.set_segments(offs)
            dev.keep_cis()
            nnz_cis = dev.nnz
            dev.distance_stats()                                 # warm-up
            (sums, counts), t_ds = wall(dev.distance_stats, dev)
            expected = expected_values("obs_exp", sums, counts, offs)
            _, t_ap = wall(lambda: dev.obs_exp_apply(expected, 1, 1), dev)
            _, t_ez = wall(dev.eliminate_zeros, dev)
            res["obs_exp_per_chromosome"] = {"entries": nnz_cis, "distance_stats_ms": t_ds * 1e3, "apply_ms": t_ap * 1e3,
                                             "eliminate_zeros_ms": t_ez * 1e3, "entries_out": dev.nnz,
                                             "distance_stats_GBps": 12.0 * nnz_cis / t_ds / 1e9,
                                             "apply_GBps": 20.0 * nnz_cis / t_ap / 1e9,
                                             "entries_per_s_whole_tool": nnz_cis / (t_ds + t_ap + t_ez)}
        kw_b = dict(kw, seed=1)
        with DeviceCSR.synthetic(bins, device=local, **kw) as a, DeviceCSR.synthetic(bins, device=local, **kw_b) as b:
            na, nb = a.nnz, b.nnz
            _, t_add = wall(lambda: a.add(b), a)
            nc = a.nnz
            res["sum"] = {"ms": t_add * 1e3, "entries_a": na, "entries_b": nb, "entries_out": nc,
                          "input_entries_per_s": (na + nb) / t_add,
                          "GBps": (2 * 12.0 * (na + nb) + 12.0 * nc) / t_add / 1e9,
                          "frac_of_peak": (2 * 12.0 * (na + nb) + 12.0 * nc) / t_add / 1e9 / peak,
                          "note": "two passes over both inputs (count, fill) + allocation of the summed store"}
        sa = hic_cpu.synth_matrix([bins[0]], **kw)
        sb = hic_cpu.synth_matrix([bins[0]], **kw_b)
        t0 = time.time()
        normalize_oracle.normalize([sa], "norm_range")
        t_cn = time.time() - t0
        t0 = time.time()
        normalize_oracle.sum_matrices([sa, sb])
        t_cs = time.time() - t0
        from oracle import transform_oracle
        small = sa[:6000, :6000].tocsr()                        # the reference's per-distance mask loop is O(n_dist x nnz)
        t0 = time.time()
        transform_oracle.obs_exp(small)
        t_ct = time.time() - t0
        res["cpu_baseline"] = {"normalize_entries_per_s": sa.nnz / t_cn, "sum_input_entries_per_s": (sa.nnz + sb.nnz) / t_cs,
                               "obs_exp_entries_per_s": small.nnz / t_ct,
                               "obs_exp_sample": "oracle/transform_oracle.py (sort-based restatement, far faster than the "
                                                 "reference's mask loop) on a 6000-bin corner (%d entries)" % small.nnz,
                               "cores": 1, "kind": "port",
                               "sample": "oracle/normalize_oracle.py (numpy/scipy restatement) on chromosome 1 of the same "
                                         "matrices (%d / %d entries)" % (sa.nnz, sb.nnz)}
        out.update({"workload": "hicNormalize norm_range / hicSumMatrices on the synthetic 10 kb matrix (%d bins, %d entries)"
                                % (int(sum(bins)), nnz)})
        out.update(res)
    elif config == "perchr5k":
        method, bins, kw = bench.synth_kwargs("perchr5k", clean=True)
        kw = dict(kw, n_trans=0)                    # cis only: block diagonal by construction
        weights = [bench.expected_nnz([b], kw["band"], kw["d0"]) for b in bins]
        owner = lpt(weights, world)
        mine = [c for c in range(len(bins)) if owner[c] == rank]
        my_bins = [bins[c] for c in mine]
        dev = DeviceCSR.synthetic(my_bins, device=local, **kw)
        offs = np.concatenate([[0], np.cumsum(my_bins)]).astype(np.int64)
        dev.set_segments(offs)
        n_local, nnz_local = dev.n, dev.nnz
        h = dev.handle
        steps = 50
        check(L.hb_dev_ice_begin(h, None))
        check(L.hb_dev_ice_loop(h, 5, 0.0, None))
        dev.sync()
        barrier()
        t0 = time.time()
        check(L.hb_dev_ice_loop(h, steps, 0.0, None))      # all chromosomes of the rank: one persistent launch
        dev.sync()
        t_loop = allmax(time.time() - t0)
        barrier()
        t0 = time.time()
        bias, iters, devn = dev.ice(500, 1e-5)
        t_conv = allmax(time.time() - t0)
        nnz_total = allsum(nnz_local)
        conv_all = allsum(float(np.all(devn < 1e-5))) == world
        dev.close()
        # KR: every chromosome is its own store; KR_THREADS host threads each run whole chromosomes on their own
        # handle / stream, so the per-mat-vec host synchronisation of one chromosome hides behind the kernels of
        # another (one tiny warm-up run first: CUDA loads kernels lazily).  Generation is outside the timed region.
        from concurrent.futures import ThreadPoolExecutor
        kr_threads = int(os.environ.get("KR_THREADS", "2"))
        with DeviceCSR.synthetic([2000], device=local, **kw) as d0:
            d0.kr()
        blocks = {c: DeviceCSR.synthetic([bins[c]], device=local, **kw) for c in mine}
        for d1 in blocks.values():
            d1.pack_values()
            d1.kr(max_outer=0)        # allocates the handle's solver workspace (device + pinned): not part of the solve
        torch.cuda.synchronize()
        barrier()
        t0 = time.time()
        order = sorted(mine, key=lambda c: -blocks[c].nnz)
        with ThreadPoolExecutor(max_workers=kr_threads) as pool:
            results = list(pool.map(lambda c: blocks[c].kr(), order))
        torch.cuda.synchronize()
        t_kr = allmax(time.time() - t0)
        mv_local = sum(r[2] + 1 for r in results)
        work = sum(float(blocks[c].nnz) * (r[2] + 1) for c, r in zip(order, results))
        for d1 in blocks.values():
            d1.close()
        out.update({
            "workload": "perchr: synthetic 5 kb human, %d chromosome blocks, %d bins, %.3g stored entries (cis only)"
                        % (len(bins), sum(bins), nnz_total),
            "ice": {"ms_per_iteration_all_chromosomes": t_loop / steps * 1e3, "nnz_per_s": nnz_total * steps / t_loop,
                    "GBps_per_gpu": 12.0 * nnz_total * steps / t_loop / world / 1e9,
                    "frac_of_peak": 12.0 * nnz_total * steps / t_loop / world / 1e9 / peak,
                    "time_to_converge_ms": t_conv * 1e3, "iterations_min_max": [int(-allmax(-float(iters.min()))), int(allmax(float(iters.max())))],
                    "all_converged": bool(conv_all)},
            "kr": {"time_all_chromosomes_ms": t_kr * 1e3, "matvec_entries_per_s": allsum(work) / t_kr,
                   "matvecs": int(allsum(mv_local)), "host_threads_per_gpu": kr_threads,
                   "includes": "KR to tol 1e-6 of every chromosome block (uint16 value layout); blocks generated and solver workspaces "
                               "allocated beforehand"},
            "communication": "none (independent chromosomes; LPT assignment)"})
    else:
        wl = {"kr1k": "kr1k", "kr25k_sharded": "kr25k", "ice10k_conv": "ice10k"}[config]
        method, bins, kw = bench.synth_kwargs(wl, clean=(wl == "ice10k"))
        n = int(sum(bins))
        t0 = time.time()
        if world == 1:
            dev = DeviceCSR.synthetic(bins, device=local, **kw)
        else:
            eq = np.linspace(0, n, world + 1).astype(np.int64)
            with DeviceCSR.synthetic(bins, row_range=(int(eq[rank]), int(eq[rank + 1])), device=local, **kw) as tmp:
                ip = np.empty(tmp.n_rows + 1, dtype=np.int64)
                check(L.hb_csr_download(tmp.handle, _lib.ptr(ip), None, None))
            counts = torch.zeros(n, dtype=torch.int64, device="cuda")
            counts[int(eq[rank]):int(eq[rank + 1])] = torch.from_numpy(np.diff(ip)).cuda()
            dist.all_reduce(counts)
            bounds = nnz_balanced_partition(np.concatenate([[0], np.cumsum(counts.cpu().numpy())]), world)
            del counts
            dev = DeviceCSR.synthetic(bins, row_range=(int(bounds[rank]), int(bounds[rank + 1])), device=local, **kw)
        keep = None
        peer_mode = False
        if world > 1:
            if os.environ.get("HB_EXCHANGE", "peer") == "peer":
                peer_mode = attach_peer(dev, rank, world)
            if not peer_mode:
                keep = attach_comm(dev, rank, world)
        torch.cuda.synchronize()
        t_gen = time.time() - t0
        nnz_total = allsum(dev.nnz)
        if method == "KR":          # warm-up: workspace allocation, lazy module load
            dev.kr()
        else:
            dev.ice(2, 0.0)
        barrier()
        t0 = time.time()
        if method == "KR":
            x, outer, mv, res = dev.kr()
            t_run = allmax(time.time() - t0)
            out.update({"kr": {"time_to_converge_ms": t_run * 1e3, "outer": outer, "matvecs": mv, "residual": res,
                               "ms_per_matvec_incl_vector_ops_and_allreduce": t_run * 1e3 / (mv + 1),
                               "nnz_per_s": nnz_total * (mv + 1) / t_run,
                               "GBps_per_gpu": (12.0 * nnz_total + 120.0 * n) * (mv + 1) / t_run / world / 1e9,
                               "frac_of_peak": (12.0 * nnz_total + 120.0 * n) * (mv + 1) / t_run / world / 1e9 / peak,
                               "x_min_max": [float(x.min()), float(x.max())]}})
        else:
            bias, iters, devn = dev.ice(500, 1e-5)
            t_run = allmax(time.time() - t0)
            out.update({"ice": {"time_to_converge_ms": t_run * 1e3, "iterations": int(iters[0]), "deviation": float(devn[0]),
                                "ms_per_iteration": t_run * 1e3 / max(int(iters[0]), 1),
                                "nnz_per_s": nnz_total * int(iters[0]) / t_run}})
        out.update({"workload": bench.workload_name(wl, bins, kw), "bins": n, "nnz": nnz_total, "generated_in_s": t_gen,
                    "communication": ("none" if world == 1 else "peer stores into every replica over NVLink (no collective)"
                                      if peer_mode else "1 NCCL sum-allreduce of %d fp64 per mat-vec/iteration" % n)})
        dev.close()
        del keep
    return out


if __name__ == "__main__":
    main()
