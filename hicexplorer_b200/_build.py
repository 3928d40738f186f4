"""Compile libhicbalance.so (sm_100a) in-tree with nvcc.  Called by
``__graft_entry__.build()``; the product never JIT-compiles at import time."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhicbalance.so")
SOURCES = ["hb_core.cu", "hb_ice.cu", "hb_filter.cu", "hb_kr.cu", "hb_synth.cu", "hb_merge.cu", "hb_xfer.cu", "hb_sym.cu", "hb_trans.cu", "hb_elem.cu", "hb_transform.cu", "hb_dense.cu", "hb_hostpack.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libhicbalance.so")
    return exe


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "hicbalance.h"),
                                                                os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=None, out=None):
    """``defines``/``out``: build a tuning variant (e.g. {"HB_SPMV_MAP": 1}) into another .so."""
    global LIB
    if out is not None:
        LIB = out
    elif not force and not _stale():
        return LIB
    objs = []
    bdir = os.path.join(PKG, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        tag = "_" + "_".join("%s%s" % kv for kv in sorted(defines.items())) if defines else ""
        obj = os.path.join(bdir, os.path.splitext(src)[0] + tag + ".o")
        dflags = ["-D%s=%s" % kv for kv in (defines or {}).items()]
        if src.endswith(".cpp"):     # host-only helpers: the host compiler directly (no -march: built here, run elsewhere)
            cxx = shutil.which("g++") or shutil.which("c++")
            if cxx is None:
                raise RuntimeError("g++ not found: cannot build libhicbalance.so")
            cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden"] + dflags + ["-c", path, "-o", obj]
        else:
            cmd = [_nvcc()] + NVCC_FLAGS + dflags + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        text = p.communicate()[0]
        log.append("== %s ==\n%s" % (src, text))
        if p.returncode != 0:
            sys.stderr.write(text)
            raise RuntimeError("nvcc failed on %s" % src)
    with open(os.path.join(bdir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.check_call(cmd)
    if out is None:
        build_io()
    return LIB


def build_io():
    """host-side file-format decoders (plain C, gcc) -> libhbio.so"""
    src = os.path.join(CSRC, "hb_io.c")
    dst = os.path.join(PKG, "libhbio.so")
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        raise RuntimeError("gcc not found: cannot build libhbio.so")
    subprocess.check_call([cc, "-O3", "-fPIC", "-shared", "-std=gnu11", "-fvisibility=hidden", "-o", dst, src, "-lz",
                           "-lpthread"])
    return dst


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
