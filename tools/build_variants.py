"""Build tuning variants of libhicbalance.so (different lane mappings / occupancy of the SpMV kernel) for
`tools/sweep.sh`; variants land in hicexplorer_b200/variants/ (git-ignored *.so, shipped to the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hicexplorer_b200 import _build  # noqa: E402

VARIANTS = {
    # round 2: how the persistent kernels read the gathered vector (HB_LOOP_GATHER: 0 weak loads = shipped,
    # 1 non-coherent loads -- NOT valid for a vector rewritten inside the launch, timing experiment only --, 2 weak + evict_last)
    "gather_weak": {"HB_LOOP_GATHER": 0},
    "gather_nc": {"HB_LOOP_GATHER": 1},
    "gather_evict_last": {"HB_LOOP_GATHER": 2},
    # round 1 sweep of the row loop (profiles/r1_spmv_variant_sweep.txt)
    "str16_c3": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3},
    "str12_c4": {"HB_SPMV_UNROLL": 12, "HB_SPMV_CTAS_PER_SM": 4},
    "str8_c4": {"HB_SPMV_UNROLL": 8, "HB_SPMV_CTAS_PER_SM": 4},
    "str24_c2": {"HB_SPMV_UNROLL": 24, "HB_SPMV_CTAS_PER_SM": 2},
}

if __name__ == "__main__":
    names = sys.argv[1:] or sorted(VARIANTS)
    out_dir = os.path.join(ROOT, "hicexplorer_b200", "variants")
    os.makedirs(out_dir, exist_ok=True)
    for name in names:
        out = os.path.join(out_dir, "libhicbalance_%s.so" % name)
        _build.build(force=True, defines=VARIANTS[name], out=out)
        log = open(os.path.join(ROOT, "hicexplorer_b200", "build", "ptxas.log")).read()
        import re
        m = re.findall(r"hb_spmv_kernelI8HbIceEpiLb1E.*?\n.*?\n.*?Used (\d+) registers", log)
        sp = re.findall(r"hb_spmv_kernelI8HbIceEpiLb1E.*?\n\s+\d+ bytes stack frame, (\d+) bytes spill stores", log)
        print(name, "regs", m, "stack/spill", sp)
