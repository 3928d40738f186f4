#!/usr/bin/env python
"""Benchmark of the matrix-correction hot path (BASELINE.json metric: ICE/KR nnz/s and iters/s vs the HBM
roofline at 1/2/4/8 B200; time-to-converge).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ice10k|kr25k|...] [--impl reference]

Headline workload (all N): BASELINE.json configs[2] -- ICE on a synthetic genome-wide 10 kb human matrix
(hg38 layout, ~309 k bins, ~1.5e9 stored entries), row-sharded in nnz-balanced blocks over the N ranks
(strong scaling: the matrix is fixed).  A "step" is one ICE iteration over the whole matrix: one fused
SpMV pass over the CSR (12 B per stored entry) + the O(n) update (+ one sum-allreduce of n doubles when
N > 1).  `value` = stored entries processed per second with the CSR resident in HBM; `e2e` = the same
through the host-buffer C-ABI (hb_csr_create from pinned host memory -> hb_ice_run(K) -> bias back),
copies inside the timed region.  configs[1] (KR, 25 kb, ~3e8 entries) is measured too at N = 1 and reported
under "kr".  `--impl reference` times the reference's CPU algorithm (numpy port in oracle/, single thread
like the reference) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
        133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
        58617616, 64444167, 46709983, 50818468, 156040895, 57227415]

WORKLOADS = {
    # name: (method, resolution bp, target stored entries, band in bins)
    "ice10k": ("ICE", 10000, 1.5e9, 16000),
    "kr25k": ("KR", 25000, 3.0e8, 8000),
    "ice25k": ("ICE", 25000, 3.0e8, 8000),
    "ice100k": ("ICE", 100000, 4.0e7, 2000),   # small case for quick checks
    "ice5k_big": ("ICE", 5000, 2.6e9, 16000),  # > 2^31 stored entries on one GPU (int64 offsets everywhere)
}


def chrom_bins(res):
    return [-(-length // res) for length in HG38]


def expected_nnz(bins, band, d0):
    tot = 0.0
    for L in bins:
        dmax = min(band, L - 1)
        d = np.arange(1, dmax + 1, dtype=np.float64)
        p = np.minimum(1.0, d0 / (d + 1.0))
        tot += L + 2.0 * np.sum(p * (L - d))
    return tot


def solve_d0(bins, band, target):
    lo, hi = 1.0, float(band)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if expected_nnz(bins, band, mid) < target:
            lo = mid
        else:
            hi = mid
    return round(0.5 * (lo + hi), 3)


def synth_kwargs(workload, clean=True):
    method, res, target, band = WORKLOADS[workload]
    bins = chrom_bins(res)
    d0 = solve_d0(bins, band, target)
    kw = dict(seed=0, band=band, d0=d0, gamma=1.0, lambda0=40.0, n_trans=32, lambda_trans=400.0,
              low_frac=0.0 if clean else 0.02, empty_frac=0.0 if clean else 0.005)
    return method, bins, kw


# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md) through NVML
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints), every 10 ms."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, device):
        self.samples = []
        self.mark = 0
        self.stop_flag = False
        self.handle = None
        self.nvml = None
        self.thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(device).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.nvml = pynvml
        except Exception as exc:
            self.error = repr(exc)

    def start(self):
        if self.nvml is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.01)

    def begin_region(self):
        self.mark = len(self.samples)

    def stop(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "error", "")]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        region = self.samples[self.mark:] or self.samples[-1:]
        try:
            mx = self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
        except Exception:
            mx = None
        bits = 0
        for _sm, rs, _pw in region:
            bits |= int(rs)
        reasons = sorted(name for bit, name in self.REASONS.items() if bits & bit)
        return {"sm_mhz": float(np.median([r[0] for r in region])) if region else None, "sm_max_mhz": mx,
                "power_w_max": max([r[2] for r in region]) if region else None, "samples": len(region),
                "reasons": reasons}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def cpu_ice_baseline(workload, seconds_budget=25.0, use_gpu_generator=False):
    """The reference's ICE loop (numpy port, oracle/ice_oracle.py == iterativeCorrection.py:40-74) on a bounded
    sample: the same generator and parameters restricted to chromosome 1."""
    from oracle import ice_oracle
    method, bins, kw = synth_kwargs(workload, clean=True)
    sample_bins = [bins[0]]
    t0 = time.time()
    if use_gpu_generator:
        from hicexplorer_b200.store import DeviceCSR
        with DeviceCSR.synthetic(sample_bins, **kw) as dev:
            m = dev.download()
    else:
        from oracle import hic_cpu
        m = hic_cpu.synth_matrix(sample_bins, **kw)
    gen_s = time.time() - t0
    nnz = m.nnz
    t0 = time.time()
    ice_oracle.ice(m.copy(), max_iter=1, tolerance=0.0)
    t1 = time.time() - t0
    k = int(max(2, min(20, (seconds_budget - t1) / max(t1 * 0.6, 1e-3))))
    t0 = time.time()
    ice_oracle.ice(m.copy(), max_iter=1 + k, tolerance=0.0)
    t2 = time.time() - t0
    per_iter = max((t2 - t1) / k, 1e-9)
    # the same loop restated in C on every host core (pthreads; the reference itself is single-threaded numpy and cannot
    # do this): an upper bound on what parallelising the reference's algorithm on this host would give
    all_cores = None
    try:
        from oracle import hic_cpu as _hc
        cores = os.cpu_count() or 1
        _hc.ice_loop_mt(m, 1, 0.0, threads=cores)
        _b, it_mt, _d, _data, secs = _hc.ice_loop_mt(m, 8, 0.0, threads=cores)
        all_cores = {"value": nnz * it_mt / secs, "unit": "nnz/s", "cores": cores, "kind": "port",
                     "what": "oracle/hic_cpu.c ice_loop_coo_mt: the reference's loop (row sums + two-factor rescale of "
                             "every entry) in C with pthreads, %d iterations on the same sample" % it_mt}
    except Exception as exc:
        all_cores = {"error": repr(exc)}
    return {"value": nnz / per_iter, "unit": "nnz/s", "cores": 1, "kind": "port", "c_port_all_cores": all_cores,
            "iters_per_s": 1.0 / per_iter, "host_cores_available": os.cpu_count(),
            "sample": "chromosome 1 of the same synthetic layout (%d bins, %d stored entries, same band/d0); "
                      "%d loop iterations of the reference's numpy ICE (oracle/ice_oracle.py), setup excluded; "
                      "generated in %.1f s" % (m.shape[0], nnz, k, gen_s)}, m


def cpu_kr_baseline(workload="kr25k", seconds_budget=20.0):
    """Restated KR (oracle/kr_oracle.py, scipy CSR mat-vec, single thread; krbalancing is not installed)."""
    from oracle import hic_cpu
    method, bins, kw = synth_kwargs(workload, clean=True)
    m = hic_cpu.synth_matrix([bins[0]], **kw)
    x = np.ones(m.shape[0])
    t0 = time.time()
    reps = 0
    while time.time() - t0 < min(seconds_budget, 8.0):
        _ = x * (m @ x)
        reps += 1
    per = (time.time() - t0) / reps
    return {"value": m.nnz / per, "unit": "nnz/s", "cores": 1, "kind": "port", "matvecs_per_s": 1.0 / per,
            "sample": "chromosome 1 of the 25 kb layout (%d bins, %d entries): %d scipy CSR mat-vecs x o (A x), "
                      "the operation inside the restated bnewt loop" % (m.shape[0], m.nnz, reps)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, _ = cpu_ice_baseline(args.workload, seconds_budget=30.0, use_gpu_generator=False)
    method, bins, kw = synth_kwargs(args.workload)
    line = {"impl": "reference", "metric": "ICE stored entries processed per second (nnz x iterations / s)",
            "value": base["value"], "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / base["iters_per_s"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, bins, kw),
                       "step": "one ICE iteration of the reference algorithm on the bounded sample described in cpu_baseline.sample"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(workload, bins, kw):
    method, res, target, band = WORKLOADS[workload]
    return ("%s synthetic genome-wide %d kb human (hg38 layout, %d bins, target %.2g stored entries, band %d, "
            "d0 %.3f, seed 0), full symmetric CSR int64/int32/fp64" % (method, res // 1000, sum(bins), target, band,
                                                                       kw["d0"]))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="ice10k", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-kr", action="store_true")
    ap.add_argument("--quick", action="store_true", help="timed ICE loop only (for ncu captures)")
    ap.add_argument("--no-all-configs", dest="all_configs", action="store_false",
                    help="skip BASELINE.json configs[3] (--perchr 5 kb) and configs[4] (KR 1 kb, 8e9 entries) after the headline")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: rows stored into peer replicas over NVLink by the SpMV epilogue (default, falls back "
                         "to nccl when peer access is unavailable) or one NCCL sum-allreduce per iteration")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    if args.quick:
        args.no_e2e = args.no_cpu = args.no_kr = True

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
    _lib.require_device()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a real (non-default) stream: the C-ABI treats a NULL stream as "the handle's own stream"
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    method, bins, kw = synth_kwargs(args.workload, clean=True)
    n = int(sum(bins))

    # ---- generate this rank's nnz-balanced row block directly in HBM ---------------------------------
    t_gen = time.time()
    if world == 1:
        dev = DeviceCSR.synthetic(bins, device=local, **kw)
        bounds = np.array([0, n])
    else:
        eq = np.linspace(0, n, world + 1).astype(np.int64)
        with DeviceCSR.synthetic(bins, row_range=(int(eq[rank]), int(eq[rank + 1])), device=local, **kw) as tmp:
            ip = np.empty(tmp.n_rows + 1, dtype=np.int64)
            check(L.hb_csr_download(tmp.handle, _lib.ptr(ip), None, None))
        counts = torch.zeros(n, dtype=torch.int64, device="cuda")
        counts[int(eq[rank]):int(eq[rank + 1])] = torch.from_numpy(np.diff(ip)).cuda()
        dist.all_reduce(counts)
        indptr_all = np.concatenate([[0], np.cumsum(counts.cpu().numpy())])
        bounds = nnz_balanced_partition(indptr_all, world)
        dev = DeviceCSR.synthetic(bins, row_range=(int(bounds[rank]), int(bounds[rank + 1])), device=local, **kw)
    peer_mode = False
    keep = None
    if world > 1:
        if args.exchange == "peer":
            peer_mode = attach_peer(dev, rank, world)
        if not peer_mode:
            keep = attach_comm(dev, rank, world)
    n_rows, n_cols, row0, nnz_local = dev.dims()
    nnz_t = torch.tensor([float(nnz_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    h = dev.handle

    exch = torch.zeros(n + 64, dtype=torch.float64, device="cuda")
    eptr = None if peer_mode else ctypes.c_void_p(exch.data_ptr())
    # One GPU, or row blocks in peer mode: K iterations = ONE cooperative launch (hb_dev_ice_loop: SpMV + update per
    # iteration separated by grid barriers, nothing on the host).  With the NCCL exchange the host has to sit between the
    # SpMV and the update, so that mode keeps one launch per operation.
    use_loop = (world == 1 or peer_mode) and os.environ.get("HB_NO_PERSISTENT") != "1"

    def ice_step():
        check(L.hb_dev_ice_spmv(h, eptr, rank, world, sptr))
        if world > 1 and not peer_mode:
            dist.all_reduce(exch[:n + world])
        check(L.hb_dev_ice_update(h, eptr, world, 0.0, sptr))

    def run_steps(k, tol=0.0):
        if use_loop:
            check(L.hb_dev_ice_loop(h, k, tol, sptr))
        else:
            for _ in range(k):
                ice_step()

    if os.environ.get("HB_BENCH_PACK") == "1":     # tuning sweeps only: time the packed layout in the main loop
        dev.pack_values()
    # ---- device-resident timing -------------------------------------------------------------------------
    check(L.hb_dev_ice_begin(h, sptr))
    run_steps(args.warmup)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if not args.quick:
        run_steps(max(args.warmup, 20))     # let the sampler spin up under load (all ranks)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.hb_launch_count()
    barrier()
    sampler.begin_region()
    ev0.record(stream)
    run_steps(args.steps)
    ev1.record(stream)
    barrier()
    launches = L.hb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = nnz_total * args.steps / (ms_total * 1e-3)
    # the dominant kernel for the roofline: the loop kernel's launch IS the timed region (K iterations per launch, so
    # bytes per launch = K x bytes per iteration); in step-level mode the SpMV launches are timed one by one
    if use_loop:
        spmv_ms = ms_step
        kernel_name = "hb_ice_loop_kernel<double> (one cooperative launch = %d iterations: SpMV + vector update)" % args.steps
    else:
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in pairs:
            a.record(stream)
            check(L.hb_dev_ice_spmv(h, eptr, rank, world, sptr))
            b.record(stream)
            if world > 1 and not peer_mode:
                dist.all_reduce(exch[:n + world])
            check(L.hb_dev_ice_update(h, eptr, world, 0.0, sptr))
        barrier()
        spmv_ms = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in pairs])))
        kernel_name = "hb_spmv_kernel<HbIceEpi,true>"
    # the same K iterations through the step-level API (one launch per operation, as in round 1) for comparison
    step_level_ms = None
    if use_loop and not args.quick:
        for _ in range(3):
            ice_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(args.steps):
            ice_step()
        s1.record(stream)
        barrier()
        step_level_ms = max_over_ranks(s0.elapsed_time(s1)) / args.steps
    # the stand-alone SpMV kernel of the step-level API next to it (N = 1): what one pass over the CSR costs without the
    # vector phases and grid barriers of the loop kernel
    spmv_alone_ms = None
    if world == 1 and not args.quick:
        check(L.hb_dev_ice_begin(h, sptr))
        for _ in range(3):
            ice_step()
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in pairs:
            a.record(stream)
            check(L.hb_dev_ice_spmv(h, eptr, rank, world, sptr))
            b.record(stream)
            check(L.hb_dev_ice_update(h, eptr, world, 0.0, sptr))
        torch.cuda.synchronize()
        spmv_alone_ms = float(np.mean([a.elapsed_time(b) for a, b in pairs]))

    # ---- time to converge (tol 1e-5, the reference default), device resident ------------------------------
    barrier()
    status = torch.zeros(4, dtype=torch.int32, pin_memory=True)
    conv_iters = 0
    converged = False
    converge_ms = None
    parity = None
    if not args.quick:
        check(L.hb_dev_ice_begin(h, sptr))
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        if use_loop:
            check(L.hb_dev_ice_loop(h, 500, 1e-5, sptr))
        else:
            for k in range(1, 501):
                check(L.hb_dev_ice_spmv(h, eptr, rank, world, sptr))
                if world > 1 and not peer_mode:
                    dist.all_reduce(exch[:n + world])
                check(L.hb_dev_ice_update(h, eptr, world, 1e-5, sptr))
                if k % 4 == 0:
                    check(L.hb_dev_ice_status(h, ctypes.c_void_p(status.data_ptr()), sptr))
                    stream.synchronize()
                    if int(status[0]):
                        break
        c1.record(stream)
        check(L.hb_dev_ice_status(h, ctypes.c_void_p(status.data_ptr()), sptr))
        stream.synchronize()
        conv_iters = int(status[2])
        converge_ms = max_over_ranks(c0.elapsed_time(c1))
        converged = bool(int(status[0])) and not bool(int(status[1]))
        check(L.hb_dev_ice_finish(h, sptr))
        stream.synchronize()
        # ---- parity record: the converged bias (and, below, the KR vector) must be BIT-identical for every N -------
        import hashlib
        bias_h = np.empty(n, dtype=np.float64)
        it_h = np.zeros(1, dtype=np.int32)
        check(L.hb_ice_results(h, _lib.ptr(bias_h), _lib.ptr(it_h), None))
        parity = {"bias_sha256": hashlib.sha256(bias_h.tobytes()).hexdigest(), "iterations": int(it_h[0]),
                  "bias_row_sum_check": None}

    # ---- KR on the SAME matrix, same sharding (north star: "10 kb ICE and KR ... on 8 x B200") -----------------
    kr10 = None
    if not args.quick and not args.no_kr and (world == 1 or peer_mode or keep is not None):
        import hashlib
        barrier()
        dev.kr()                      # first call: workspace allocation, lazy module load
        barrier()
        t0 = time.time()
        xk, kouter, kmv, kres = dev.kr()
        torch.cuda.synchronize()
        kr_s = max_over_ranks(time.time() - t0)
        peak_k, _ = measured_peak()
        bytes_mv = 12.0 * nnz_total + 120.0 * n
        per_mv_ms = kr_s * 1e3 / max(kmv + 1, 1)
        kr10 = {"what": "hb_kr_run to tol 1e-6 on the %s matrix of this run, %d row block(s): ONE cooperative launch "
                        "(hb_kr_solve_kernel) per rank, host-timed call (x copied back inside)" % (args.workload, world),
                "time_to_converge_ms": kr_s * 1e3, "outer_iterations": int(kouter), "matvecs": int(kmv),
                "residual": float(kres), "ms_per_matvec_incl_vector_work": per_mv_ms,
                "GBps_per_gpu": bytes_mv / world / (per_mv_ms * 1e-3) / 1e9,
                "frac_of_peak_per_gpu": bytes_mv / world / (per_mv_ms * 1e-3) / 1e9 / peak_k,
                "frac_of_nominal_8TBps": bytes_mv / world / (per_mv_ms * 1e-3) / 1e9 / 8000.0,
                "nnz_per_s": nnz_total * (kmv + 1) / kr_s}
        if parity is not None:
            parity["kr_x_sha256"] = hashlib.sha256(np.ascontiguousarray(xk).tobytes()).hexdigest()
            parity["kr_matvecs"] = int(kmv)
            parity["kr_outer"] = int(kouter)

    # ---- N > 1: rank 0 repeats both solves on ONE GPU with the whole matrix and compares the digests ----------
    if parity is not None and world > 1:
        ok_flag = torch.ones(1, dtype=torch.int32, device="cuda")
        if rank == 0:
            try:
                import hashlib
                with DeviceCSR.synthetic(bins, device=local, **kw) as one:
                    b1, it1, _d1 = one.ice(500, 1e-5)
                    single = {"bias_sha256": hashlib.sha256(np.ascontiguousarray(b1).tobytes()).hexdigest(),
                              "iterations": int(it1[0])}
                    if "kr_x_sha256" in parity:
                        x1, o1, mv1, _r1 = one.kr()
                        single.update(kr_x_sha256=hashlib.sha256(np.ascontiguousarray(x1).tobytes()).hexdigest(),
                                      kr_matvecs=int(mv1), kr_outer=int(o1))
                parity["single_gpu_same_run"] = single
                parity["identical_to_single_gpu"] = all(parity.get(k) == v for k, v in single.items())
            except Exception as exc:
                parity["single_gpu_same_run"] = {"error": repr(exc)}
                parity["identical_to_single_gpu"] = None
            if parity.get("identical_to_single_gpu") is False:
                ok_flag[0] = 0
        dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN)
        if int(ok_flag.item()) != 1:
            raise SystemExit("PARITY FAILURE: the sharded result differs from the single-GPU result: %r" % (parity,))

    # ---- optional lossless narrow value layout (hb_csr_pack_values): same results bit for bit, fewer bytes streamed
    packed = None
    if not args.quick:
        kind = dev.pack_values()
        if kind:
            vbytes = {1: 2, 2: 4}[kind]
            check(L.hb_dev_ice_begin(h, sptr))
            run_steps(5)
            barrier()
            pk0, pk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            psteps = min(args.steps, 100)
            pk0.record(stream)
            run_steps(psteps)
            pk1.record(stream)
            barrier()
            pms = max_over_ranks(pk0.elapsed_time(pk1)) / psteps
            packed = {"layout": "int32 column id + %s value (lossless copy of the fp64 counts; fp64 stays the store of "
                                "record)" % {1: "uint16", 2: "float32"}[kind],
                      "bytes_per_entry": 4 + vbytes, "ms_per_step": pms, "nnz_per_s": nnz_total / (pms * 1e-3),
                      "GBps_per_gpu": (4.0 + vbytes) * nnz_total / world / (pms * 1e-3) / 1e9}

    # ---- end to end through the host-buffer C-ABI ------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            ip_h = torch.empty(n_rows + 1, dtype=torch.int64, pin_memory=True)
            ix_h = torch.empty(max(nnz_local, 1), dtype=torch.int32, pin_memory=True)
            da_h = torch.empty(max(nnz_local, 1), dtype=torch.float64, pin_memory=True)
            check(L.hb_csr_download(h, ctypes.c_void_p(ip_h.data_ptr()), ctypes.c_void_p(ix_h.data_ptr()),
                                    ctypes.c_void_p(da_h.data_ptr())))
            bias_h = torch.empty(n, dtype=torch.float64, pin_memory=True)
            iters_h = np.zeros(1, dtype=np.int32)
            devv = np.zeros(1, dtype=np.float64)
            dev.close()
            dev = None
            # one-off per process, like the CUDA context: the upload path pins its staging buffers (~0.8 ms per MB) the
            # first time it runs.  A small block of the same matrix goes through hb_csr_create once before the clock starts.
            t_w = time.time()
            r_w = int(min(n_rows, max(1, np.searchsorted(ip_h.numpy(), int(ip_h[0]) + min(nnz_local, 48_000_000)))))
            h_w = ctypes.c_void_p()
            check(L.hb_csr_create(ctypes.byref(h_w), r_w, n_cols, row0, int(ip_h[r_w] - ip_h[0]), ctypes.c_void_p(ip_h.data_ptr()),
                                  ctypes.c_void_p(ix_h.data_ptr()), 0, ctypes.c_void_p(da_h.data_ptr()), local))
            check(L.hb_csr_destroy(h_w))
            t_w = time.time() - t_w
            barrier()
            t0 = time.time()
            h2 = ctypes.c_void_p()
            check(L.hb_csr_create(ctypes.byref(h2), n_rows, n_cols, row0, nnz_local, ctypes.c_void_p(ip_h.data_ptr()),
                                  ctypes.c_void_p(ix_h.data_ptr()), 0, ctypes.c_void_p(da_h.data_ptr()), local))
            t_create = time.time() - t0
            ingest_kind, ingest_bytes = ctypes.c_int(0), ctypes.c_int64(0)
            check(L.hb_csr_ingest_info(h2, ctypes.byref(ingest_kind), ctypes.byref(ingest_bytes)))
            dev2 = DeviceCSR(h2, local)
            keep2 = None
            if world > 1 and not (peer_mode and attach_peer(dev2, rank, world)):
                keep2 = attach_comm(dev2, rank, world)
            t_attach = time.time() - t0 - t_create
            check(L.hb_ice_run(h2, args.steps, 0.0, ctypes.c_void_p(bias_h.data_ptr()), _lib.ptr(iters_h),
                               _lib.ptr(devv)))
            t_run = time.time() - t0 - t_create - t_attach
            barrier()
            e2e_s = max_over_ranks(time.time() - t0)
            h2d_total = sum_over_ranks(float(ingest_bytes.value))
            breakdown = {"hb_csr_create_s": max_over_ranks(t_create), "attach_exchange_s": max_over_ranks(t_attach),
                         "hb_ice_run_s": max_over_ranks(t_run),
                         "host_bytes_read_per_gpu": 12.0 * nnz_local + 8.0 * (n_rows + 1),
                         "pcie_bytes_per_gpu": float(ingest_bytes.value),
                         "value_layout_after_create": {0: "fp64 only", 1: "fp64 + uint16 (narrow ingest: counts converted to uint16 on "
                                                       "the host cores, 2 B per entry over PCIe; the chunks the cores did not get to "
                                                       "went as fp64 and were narrowed on the device)", 2: "fp64 + float"}[ingest_kind.value],
                         "untimed_first_use_s": max_over_ranks(t_w),
                         "host_GBps_per_gpu_inside_create": 12.0 * nnz_local / max(t_create, 1e-9) / 1e9,
                         "note": "rank-wise maxima of host timestamps; the upload dominates hb_csr_create (pinned host memory, "
                                 "PCIe Gen5 x16 per GPU; with N ranks the host's cores and memory bandwidth are shared)"}
            dev2.close()
            del keep2
            # the same call sequence run to the reference's own stopping rule (tolerance 1e-5): the whole job a caller of
            # iterativeCorrection() waits for, upload and read-back included
            barrier()
            t0 = time.time()
            h3 = ctypes.c_void_p()
            check(L.hb_csr_create(ctypes.byref(h3), n_rows, n_cols, row0, nnz_local, ctypes.c_void_p(ip_h.data_ptr()),
                                  ctypes.c_void_p(ix_h.data_ptr()), 0, ctypes.c_void_p(da_h.data_ptr()), local))
            dev3 = DeviceCSR(h3, local)
            keep3 = None
            if world > 1 and not (peer_mode and attach_peer(dev3, rank, world)):
                keep3 = attach_comm(dev3, rank, world)
            iters_c = np.zeros(1, dtype=np.int32)
            check(L.hb_ice_run(h3, 500, 1e-5, ctypes.c_void_p(bias_h.data_ptr()), _lib.ptr(iters_c), _lib.ptr(devv)))
            barrier()
            conv_s = max_over_ranks(time.time() - t0)
            dev3.close()
            del keep3
            # and the strictest reading of "the call a user makes": the Python drop-in itself,
            # iterativeCorrection(scipy_matrix, M=steps) -> (corrected scipy CSR, bias), single GPU only: canonical-form and
            # symmetry checks, upload, `steps` iterations, emission of the corrected matrix and its way back into a
            # scipy matrix (11.5 GB D2H) are all inside the timed region
            dropin = None
            if world == 1 and not os.environ.get("HB_BENCH_SKIP_DROPIN"):
                try:
                    from scipy import sparse as _sp
                    from hicexplorer_b200.iterativeCorrection import iterativeCorrection as _ice
                    m_host = _sp.csr_matrix((da_h.numpy(), ix_h.numpy(), ip_h.numpy()), shape=(n, n))
                    t0 = time.time()
                    w_host, b_host = _ice(m_host, M=args.steps, tolerance=0.0, device=local)
                    drop_s = time.time() - t0
                    dropin = {"seconds": drop_s, "nnz_per_s": nnz_total * args.steps / drop_s, "iterations": args.steps,
                              "d2h_bytes": 8.0 * w_host.nnz + 8.0 * n,
                              "what": "hicexplorer_b200.iterativeCorrection.iterativeCorrection(scipy CSR in pinned host "
                                      "memory, M=steps, tolerance=0) -> (corrected scipy CSR, bias)"}
                    del w_host, b_host, m_host
                except Exception as exc:
                    dropin = {"error": repr(exc)}
            e2e = {"value": nnz_total * args.steps / e2e_s, "unit": "nnz/s",
                   "h2d_bytes_per_step": h2d_total / args.steps,
                   "d2h_bytes_per_step": 8.0 * n * world / args.steps, "seconds": e2e_s, "iterations": int(iters_h[0]),
                   "what": "hb_csr_create (pinned host CSR int64/int32/fp64 -> HBM; h2d bytes are what crossed PCIe, as "
                           "reported by hb_csr_ingest_info) + hb_ice_run(max_iter=steps, tol=0) + bias to host; "
                           "one call = `steps` iterations, so the one-off copy is amortised over them; the library's first upload "
                           "in a process (pinning of its staging buffers) is done on a 48 M-entry block before the clock starts "
                           "(breakdown.untimed_first_use_s)",
                   "breakdown": breakdown,
                   "to_convergence": {"seconds": conv_s, "iterations": int(iters_c[0]), "tolerance": 1e-5,
                                      "nnz_per_s": nnz_total * int(iters_c[0]) / conv_s,
                                      "what": "same calls with max_iter=500, tol=1e-5: upload + every iteration the "
                                              "reference's stopping rule needs + bias to host"},
                   "python_dropin": dropin}
        except Exception as exc:  # keep the device-resident numbers if the host leg cannot run (e.g. host RAM)
            e2e = {"value": None, "unit": "nnz/s", "error": repr(exc)}
    if dev is not None:
        dev.close()
    ip_h = ix_h = da_h = None   # release the pinned copies before the next leg

    # ---- KR (configs[1]) at N = 1 ------------------------------------------------------------------------------
    kr = None
    if world == 1 and not args.no_kr:
        kr = bench_kr(torch, L, check, _lib, DeviceCSR, local, stream, sptr)

    # ---- BASELINE.json configs[3] and configs[4] on the same N GPUs (tools/bench_configs.py) ---------------------
    # --perchr ICE + KR on the 5 kb matrix (24 chromosomes, no communication) and KR on the genome-wide 1 kb matrix
    # (3.09 M bins, 8.0e9 entries = 96 GB of CSR, row blocks over the N GPUs)
    other_configs = None
    if args.all_configs and not args.quick:
        other_configs = {}
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        torch.cuda.set_stream(torch.cuda.default_stream())
        for name in ("perchr5k", "kr1k"):
            try:
                barrier()
                torch.cuda.empty_cache()
                other_configs[name] = bench_configs.run(name)
            except Exception as exc:
                other_configs[name] = {"error": repr(exc)}
        torch.cuda.set_stream(stream)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    traffic = None
    traffic_note = None
    try:
        # DRAM bytes of the same kernel on the same workload from this round's `ncu --set full` capture
        # (profiles/r2_traffic.json, written by tools/ncu_traffic.py next to the raw export); per ITERATION, like
        # `achieved`.  The entry is only used while the kernel sources it was measured on are unchanged.
        import hashlib
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(args.workload)
        src = b"".join(open(os.path.join(ROOT, "hicexplorer_b200", "csrc", f), "rb").read()
                       for f in ("hb_spmv.cuh", "hb_ice.cu", "hb_internal.cuh"))
        if tr and tr["nnz"] == nnz_total and world == 1:
            if tr.get("source_sha256") == hashlib.sha256(src).hexdigest():
                traffic = float(tr["dram_bytes_read"] + tr["dram_bytes_write"]) / float(tr.get("iterations", 1))
            else:
                traffic_note = "profiles/r2_traffic.json was captured on older kernel sources: not reported"
    except Exception:
        pass
    bytes_spmv = 12.0 * nnz_total + 24.0 * n
    achieved = bytes_spmv / world / (spmv_ms * 1e-3) / 1e9   # per GPU: each rank streams its own share
    line = {
        "metric": "ICE stored entries processed per second (nnz x iterations / s)",
        "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "iters_per_s": 1e3 / ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, bins, kw), "bins": n, "nnz": nnz_total,
                   "parallelism": ("%d nnz-balanced row block(s); " % world) + (
                       "rows stored into every rank's replica over NVLink by the SpMV epilogue (CUDA IPC peer stores + "
                       "arrival flags), no collective and no host in the loop: ONE persistent cooperative launch per rank "
                       "runs all timed iterations" if peer_mode else
                       "1 NCCL sum-allreduce of n fp64 per iteration" if world > 1 else "single GPU"),
                   "l2": "inputs larger than L2 (%.1f GB of CSR per GPU vs 126 MB)" % (12e-9 * nnz_total / world),
                   "generated_in_s": round(t_gen, 2)},
        "time_to_converge": {"ms": converge_ms, "iterations": conv_iters, "converged": converged, "tolerance": 1e-5},
        "step_level_ms_per_step": step_level_ms,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kernel_name, "peak_source": peak_src,
                     "bytes_per_iteration": bytes_spmv / world, "avg_iteration_ms": spmv_ms,
                     "launch": ("one launch = %d iterations; bytes and time are per iteration" % args.steps) if use_loop
                     else "one launch = one iteration's SpMV",
                     "standalone_spmv_kernel_ms": spmv_alone_ms,
                     "standalone_spmv_GBps": (bytes_spmv / (spmv_alone_ms * 1e-3) / 1e9) if spmv_alone_ms else None,
                     "traffic_note": traffic_note,
                     "frac_of_nominal_8TBps": achieved / 8000.0},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if packed is not None:
        line["packed_values"] = packed
    if parity is not None:
        line["parity"] = parity
    if kr10 is not None:
        line["kr_same_matrix"] = kr10
    if other_configs is not None:
        line["other_configs"] = other_configs
    if kr is not None:
        line["kr"] = kr
    if not args.no_cpu and world == 1:
        base, _ = cpu_ice_baseline(args.workload, seconds_budget=25.0, use_gpu_generator=True)
        line["cpu_baseline"] = base
        if kr is not None:
            line["kr"]["cpu_baseline"] = cpu_kr_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_kr(torch, L, check, _lib, DeviceCSR, local, stream, sptr):
    """configs[1]: KR on the synthetic 25 kb matrix, 1 GPU: mat-vec throughput + time to converge."""
    method, bins, kw = synth_kwargs("kr25k", clean=False)
    n = int(sum(bins))
    with DeviceCSR.synthetic(bins, device=local, **kw) as dev:
        nnz = dev.nnz
        h = dev.handle
        x = torch.ones(n, dtype=torch.float64, device="cuda")
        v = torch.ones(n, dtype=torch.float64, device="cuda")
        p = torch.ones(n, dtype=torch.float64, device="cuda")
        w = torch.zeros(n, dtype=torch.float64, device="cuda")
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

        def mv():
            check(L.hb_dev_spmv(h, P(x), 1e-5, P(x), P(v), P(p), P(w), 0, sptr))
        for _ in range(5):
            mv()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        e0.record(stream)
        for _ in range(reps):
            mv()
        e1.record(stream)
        torch.cuda.synchronize()
        mv_ms = e0.elapsed_time(e1) / reps
        t0 = time.time()
        xs, outer, matvecs, resid = dev.kr()
        kr_s = time.time() - t0
    peak, _src = measured_peak()
    bytes_mv = 12.0 * nnz + 120.0 * n
    return {"workload": workload_name("kr25k", bins, kw) + " incl. 2% low-coverage and 0.5% empty bins",
            "bins": n, "nnz": nnz, "matvec_ms": mv_ms, "matvec_nnz_per_s": nnz / (mv_ms * 1e-3),
            "matvec_GBps": bytes_mv / (mv_ms * 1e-3) / 1e9, "matvec_frac_of_peak": bytes_mv / (mv_ms * 1e-3) / 1e9 / peak,
            "time_to_converge_ms": kr_s * 1e3, "outer_iterations": outer, "matvecs": matvecs, "residual": resid,
            "nnz_per_s_converge": nnz * matvecs / kr_s}


if __name__ == "__main__":
    main()
