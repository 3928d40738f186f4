#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on N >= 2 GPUs:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multigpu_check.py
Compares the sharded pipelines (real NCCL all-reduces between processes) with the single-GPU pipelines and the CLI
run under torchrun with the same command on one GPU.  Exit code 0 = all checks passed."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from scipy import sparse

    from hicexplorer_b200.pipeline import ice_pipeline, ice_pipeline_sharded, kr_pipeline, kr_pipeline_sharded
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    z = np.load(os.path.join(ROOT, "tests", "golden", "gm12878.npz"))
    n = int(z["nbins"])
    up = sparse.csr_matrix((z["count"].astype(np.float64), (z["bin1"], z["bin2"])), shape=(n, n))
    full = sparse.csr_matrix(up + sparse.triu(up, 1).T)
    full.sort_indices()
    ok = True

    def check(name, cond):
        nonlocal ok
        if rank == 0:
            print(("PASS " if cond else "FAIL ") + name, flush=True)
        ok = ok and bool(cond)

    # the host-collective exchange first (one NCCL allreduce per SpMV, step-level kernels) ...
    os.environ["HB_EXCHANGE"] = "nccl"
    got_n = ice_pipeline_sharded(full, (-1.5, 5.0), max_iter=500)
    kr_n = kr_pipeline_sharded(full, rescale=True)
    # ... then the default: peer stores from the SpMV epilogue inside ONE persistent launch per rank
    os.environ["HB_EXCHANGE"] = "peer"
    got = ice_pipeline_sharded(full, (-1.5, 5.0), max_iter=500)
    if rank == 0:
        ref = ice_pipeline(full, (-1.5, 5.0), max_iter=500, device=local)
        check("ICE sharded, NCCL exchange: bias bit-identical to 1 GPU", np.array_equal(got_n["bias"], ref["bias"]) and
              int(got_n["iterations"][0]) == int(ref["iterations"][0]))
        check("ICE sharded: masks bit-exact", np.array_equal(got["zero_bins"], ref["zero_bins"]) and
              np.array_equal(got["outliers_rel"], ref["outliers_rel"]))
        check("ICE sharded: bias bit-identical to 1 GPU", np.array_equal(got["bias"], ref["bias"]))
        check("ICE sharded: iterations equal (%d)" % int(ref["iterations"][0]), int(got["iterations"][0]) == int(ref["iterations"][0]))
        check("ICE sharded: corrected matrix pattern + values", np.array_equal(got["matrix"].indices, ref["matrix"].indices) and
              np.allclose(got["matrix"].data, ref["matrix"].data, rtol=1e-12, atol=0))
        check("ICE vs reference run (bias 1e-9)", np.allclose(got["bias"], z["ice_bias"], rtol=1e-9, atol=0))
    offs = np.concatenate([[0], np.flatnonzero(np.diff(np.searchsorted(z["chrom_offset"], np.arange(n), side="right"))) + 1, [n]])
    gotp = ice_pipeline_sharded(full, (-1.5, 5.0), max_iter=300, chr_offsets=offs, perchr=True)
    if rank == 0:
        refp = ice_pipeline(full, (-1.5, 5.0), max_iter=300, chr_offsets=offs, perchr=True, device=local)
        check("perchr ICE sharded: masks + bias identical", np.array_equal(gotp["outliers_rel"], refp["outliers_rel"]) and
              np.array_equal(gotp["bias"], refp["bias"]) and np.array_equal(gotp["iterations"], refp["iterations"]))
        check("perchr ICE sharded: matrix identical", gotp["matrix"].nnz == refp["matrix"].nnz and
              np.allclose(gotp["matrix"].data, refp["matrix"].data, rtol=1e-12, atol=0))
    # peer-store exchange (NVLink P2P stores from the SpMV epilogue instead of the NCCL callback)
    from hicexplorer_b200.dist import attach_peer, nnz_balanced_partition
    from hicexplorer_b200.store import DeviceCSR
    ice_in = None
    if True:
        from oracle import ice_oracle
        ice_in = ice_oracle.correct_ice_pipeline(full, -1.5, 5.0, max_iter=1)["ice_input"]
        ice_in.sort_indices()
        bounds = nnz_balanced_partition(ice_in.indptr, world)
        with DeviceCSR.from_scipy(ice_in, device=local, row_range=(int(bounds[rank]), int(bounds[rank + 1]))) as dv:
            peer_ok = attach_peer(dv, rank, world)
            check("peer exchange: attach (CUDA IPC + P2P)", peer_ok)
            if peer_ok:
                pb, pit, pdev = dv.ice(500, 1e-5)
                px, pouter, pmv, pres = dv.kr()
                if rank == 0:
                    with DeviceCSR.from_scipy(ice_in, device=local) as d1:
                        b1, it1, _ = d1.ice(500, 1e-5)
                        x1, o1, mv1, _ = d1.kr()
                    check("peer exchange: ICE bias bit-identical to 1 GPU, %d iterations" % int(it1[0]),
                          np.array_equal(pb, b1) and int(pit[0]) == int(it1[0]))
                    check("peer exchange: KR x bit-identical to 1 GPU, %d mat-vecs" % mv1, np.array_equal(px, x1) and pmv == mv1)
        dist.barrier()
    # the reference's tolerant symmetry test (iterativeCorrection.py:35) on row blocks: exact ratio through the block exchange
    from hicexplorer_b200.dist import attach_comm
    for delta, accept in ((1e-3, True), (25.0, False)):
        skew = ice_in.copy().tolil()
        i0, j0 = 3, int(ice_in[3].indices[ice_in[3].indices > 3][0])
        skew[i0, j0] = skew[i0, j0] + delta                     # one entry differs from its mirror image
        skew = sparse.csr_matrix(skew)
        skew.sort_indices()
        bounds = nnz_balanced_partition(skew.indptr, world)
        with DeviceCSR.from_scipy(skew, device=local, row_range=(int(bounds[rank]), int(bounds[rank + 1]))) as dv:
            keep_cb = attach_comm(dv, rank, world)
            ratio_n = dv.symmetry_ratio_sharded(dist, rank, world)
            del keep_cb
        if rank == 0:
            with DeviceCSR.from_scipy(skew, device=local) as d1:
                ratio_1 = d1.symmetry_ratio()
            want = abs(skew - skew.T).mean() / abs(skew.mean())
            check("asymmetric input (delta %g): sharded ratio %.3e == single-GPU ratio == numpy, %s like the reference"
                  % (delta, ratio_n, "accepted" if accept else "rejected"),
                  abs(ratio_n / ratio_1 - 1) < 1e-9 and abs(ratio_n / want - 1) < 1e-9 and (ratio_n <= 1e-10) == accept)
        dist.barrier()
    kr = kr_pipeline_sharded(full, rescale=True)
    if rank == 0:
        kref = kr_pipeline(full, rescale=True, want_matrix=False, device=local)
        check("KR sharded, NCCL exchange: x bit-identical, same mat-vec count", np.array_equal(kr_n["x"], kref["x"]) and
              kr_n["matvecs"] == kref["matvecs"])
        check("KR sharded: x bit-identical, same mat-vec count", np.array_equal(kr["x"], kref["x"]) and kr["matvecs"] == kref["matvecs"])
        check("KR sharded: rescale factor identical", kr["factor"] == kref["factor"])
    # the CLI under torchrun
    from hicexplorer_b200.hicCorrectMatrix import main as cli
    from hicexplorer_b200.hicmatrix_lite import hiCMatrix
    tmp = tempfile.gettempdir()
    src = os.path.join(tmp, "mg_in.cool")
    if rank == 0:
        ma = hiCMatrix()
        ma.matrix = full
        off = z["chrom_offset"]
        ma.cut_intervals = [(str(c + 1), int((i - off[c]) * 2500000), int((i - off[c] + 1) * 2500000), 1.0)
                            for c in range(len(off) - 1) for i in range(off[c], off[c + 1])]
        ma._index_intervals()
        ma.save(src)
    dist.barrier()          # the CLI reuses this process group
    out = os.path.join(tmp, "mg_out.cool")
    cli(["correct", "-m", src, "-o", out, "--correctionMethod", "ICE", "--filterThreshold", "-1.5", "5"])
    if rank == 0:
        new = hiCMatrix(out)
        cf = np.asarray(new.correction_factors).ravel()
        check("CLI under torchrun: factors match the single-GPU pipeline", np.array_equal(cf[ref["kept"]], ref["bias"]) and
              np.isnan(cf).sum() == n - len(ref["kept"]))
        from hicexplorer_b200.pipeline import scatter_to_frame
        want = sparse.triu(scatter_to_frame(ref["matrix"], ref["kept"], n), 0, format="csr")
        want.eliminate_zeros()
        want.sort_indices()
        got_up = sparse.triu(new.matrix, 0, format="csr")
        got_up.sort_indices()
        check("CLI under torchrun: matrix written from per-rank file-layout blocks equals the single-GPU result",
              np.array_equal(got_up.indptr, want.indptr) and np.array_equal(got_up.indices, want.indices) and
              np.allclose(got_up.data, want.data, rtol=1e-12, atol=0))
    # genome-wide KR to .cool under torchrun: row blocks built on the device from the stored upper triangle
    outkc = os.path.join(tmp, "mg_kr.cool")
    cli(["correct", "-m", src, "-o", outkc, "--correctionMethod", "KR"])
    if rank == 0:
        kc = hiCMatrix(outkc)
        kr1 = kr_pipeline(full, rescale=True, want_matrix=False, device=local)
        check("CLI KR under torchrun (.cool): factors bit-identical to 1 GPU, counts untouched",
              np.array_equal(np.asarray(kc.correction_factors).ravel(), kr1["x"]) and abs(kc.matrix - full).sum() == 0)
    # a chromosome subset in file order, selected on the device on every rank before the rows are dealt out
    outs = os.path.join(tmp, "mg_sub.h5")
    cli(["correct", "-m", src, "-o", outs, "--correctionMethod", "ICE", "--filterThreshold", "-1.5", "5", "--chromosomes",
         "2", "3", "7"])
    if rank == 0:
        os.environ["WORLD_SIZE"] = "1"
        outs1 = os.path.join(tmp, "mg_sub_1gpu.h5")
        cli(["correct", "-m", src, "-o", outs1, "--correctionMethod", "ICE", "--filterThreshold", "-1.5", "5", "--chromosomes",
             "2", "3", "7", "--device", str(local)])
        os.environ["WORLD_SIZE"] = str(world)
        a, b = hiCMatrix(outs), hiCMatrix(outs1)
        check("CLI --chromosomes subset under torchrun: file equals the single-GPU run",
              np.array_equal(np.isnan(np.asarray(a.correction_factors)), np.isnan(np.asarray(b.correction_factors))) and
              np.array_equal(np.nan_to_num(np.asarray(a.correction_factors)), np.nan_to_num(np.asarray(b.correction_factors))) and
              (a.matrix != b.matrix).nnz == 0)
    dist.barrier()

    def same_file(a_path, b_path):
        a, b = hiCMatrix(a_path), hiCMatrix(b_path)
        fa, fb = np.asarray(a.correction_factors).ravel(), np.asarray(b.correction_factors).ravel()
        return (np.array_equal(np.isnan(fa), np.isnan(fb)) and np.array_equal(np.nan_to_num(fa), np.nan_to_num(fb)) and
                a.matrix.shape == b.matrix.shape and (a.matrix != b.matrix).nnz == 0 and
                np.array_equal(np.asarray(a.nan_bins), np.asarray(b.nan_bins)))

    def both_ways(name, argv_tail, suffix):
        """the same command under torchrun (all ranks) and on one GPU (rank 0): the files must be identical"""
        out_n = os.path.join(tmp, "mg_%s_n%s" % (name, suffix))
        cli(["correct", "-m", src, "-o", out_n] + argv_tail)
        if rank == 0:
            os.environ["WORLD_SIZE"] = "1"
            out_1 = os.path.join(tmp, "mg_%s_1%s" % (name, suffix))
            cli(["correct", "-m", src, "-o", out_1] + argv_tail + ["--device", str(local)])
            os.environ["WORLD_SIZE"] = str(world)
            check("CLI %s under torchrun: file equals the single-GPU run" % name, same_file(out_n, out_1))
        dist.barrier()

    # genome-wide KR to .h5: every rank emits its rows of the balanced matrix (hb_kr_scaled_upper on a row block)
    both_ways("KR_h5", ["--correctionMethod", "KR"], ".h5")
    both_ways("KR_h5_subset", ["--correctionMethod", "KR", "--chromosomes", "2", "3", "7"], ".h5")
    # the optional ICE steps on row blocks: trans percentile (8 summed histogram passes), clipping, inflation mask,
    # coverage filter
    both_ways("ICE_transcut_inflation", ["--correctionMethod", "ICE", "--filterThreshold", "-1.5", "5", "--transCutoff", "5",
                                         "--clipTrans", "--inflationCutoff", "3", "--sequencedCountCutoff", "0.5"], ".h5")
    both_ways("ICE_skipdiag_perchr", ["--correctionMethod", "ICE", "--filterThreshold", "-1.5", "5", "--perchr",
                                      "--skipDiagonal"], ".h5")
    # --perchr KR: whole chromosomes dealt to the ranks (LPT), no communication; rank 0 assembles the file
    outk = os.path.join(tmp, "mg_kr_perchr.h5")
    cli(["correct", "-m", src, "-o", outk, "--correctionMethod", "KR", "--perchr"])
    if rank == 0:
        os.environ["WORLD_SIZE"] = "1"
        out1 = os.path.join(tmp, "mg_kr_perchr_1gpu.h5")
        cli(["correct", "-m", src, "-o", out1, "--correctionMethod", "KR", "--perchr", "--device", str(local)])
        a, b = hiCMatrix(outk), hiCMatrix(out1)
        check("CLI --perchr KR under torchrun: file equals the single-GPU run",
              np.array_equal(np.asarray(a.correction_factors), np.asarray(b.correction_factors)) and
              (a.matrix != b.matrix).nnz == 0)
        print("ALL PASS" if ok else "SOME FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
