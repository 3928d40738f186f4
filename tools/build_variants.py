"""Build tuning variants of libhicbalance.so (different lane mappings / occupancy of the SpMV kernel) for
`tools/sweep.sh`; variants land in hicexplorer_b200/variants/ (git-ignored *.so, shipped to the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hicexplorer_b200 import _build  # noqa: E402

VARIANTS = {
    "pk_u16_c3": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3},
    "pk_u24_c3": {"HB_SPMV_UNROLL": 24, "HB_SPMV_CTAS_PER_SM": 3},
    "pk_u32_c2": {"HB_SPMV_UNROLL": 32, "HB_SPMV_CTAS_PER_SM": 2},
    "pk_u20_c3": {"HB_SPMV_UNROLL": 20, "HB_SPMV_CTAS_PER_SM": 3},
    "pk_u16_c4": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 4},
    "pk_u12_c4": {"HB_SPMV_UNROLL": 12, "HB_SPMV_CTAS_PER_SM": 4},
    "pk_u12_c5": {"HB_SPMV_UNROLL": 12, "HB_SPMV_CTAS_PER_SM": 5},
    "pk_u8_c6": {"HB_SPMV_UNROLL": 8, "HB_SPMV_CTAS_PER_SM": 6},
    "str8_c4": {"HB_SPMV_UNROLL": 8, "HB_SPMV_CTAS_PER_SM": 4},
    "str8_c5": {"HB_SPMV_UNROLL": 8, "HB_SPMV_CTAS_PER_SM": 5},
    "str4_c6": {"HB_SPMV_UNROLL": 4, "HB_SPMV_CTAS_PER_SM": 6},
    "str4_c8": {"HB_SPMV_UNROLL": 4, "HB_SPMV_CTAS_PER_SM": 8},
    "str16_c3": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3},
    "str8_c4_g8": {"HB_SPMV_UNROLL": 8, "HB_SPMV_CTAS_PER_SM": 4, "HB_ROWS_PER_GRAB": 8},
    "str16_c2": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 2},
    "str16_c4": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 4},
    "str24_c2": {"HB_SPMV_UNROLL": 24, "HB_SPMV_CTAS_PER_SM": 2},
    "str32_c2": {"HB_SPMV_UNROLL": 32, "HB_SPMV_CTAS_PER_SM": 2},
    "str12_c3": {"HB_SPMV_UNROLL": 12, "HB_SPMV_CTAS_PER_SM": 3},
    "str16_c3_h1": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3, "HB_STREAM_HINT": 1},
    "str16_c3_h2": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3, "HB_STREAM_HINT": 2},
    "str16_c3_h3": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3, "HB_STREAM_HINT": 3},
    "str16_c3_g2": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3, "HB_ROWS_PER_GRAB": 2},
    "str16_c3_g1": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3, "HB_ROWS_PER_GRAB": 1},
    "str12_c3_g2": {"HB_SPMV_UNROLL": 12, "HB_SPMV_CTAS_PER_SM": 3, "HB_ROWS_PER_GRAB": 2},
    "str12_c4_g2": {"HB_SPMV_UNROLL": 12, "HB_SPMV_CTAS_PER_SM": 4, "HB_ROWS_PER_GRAB": 2},
    "str16_c3_g2_h2": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 3, "HB_ROWS_PER_GRAB": 2, "HB_STREAM_HINT": 2},
    "str20_c3_g2": {"HB_SPMV_UNROLL": 20, "HB_SPMV_CTAS_PER_SM": 3, "HB_ROWS_PER_GRAB": 2},
    "str16_t128c6_g2": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 6, "HB_SPMV_THREADS": 128, "HB_ROWS_PER_GRAB": 2},
    "str16_t128c6": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 6, "HB_SPMV_THREADS": 128},
    "str16_t512c1": {"HB_SPMV_UNROLL": 16, "HB_SPMV_CTAS_PER_SM": 1, "HB_SPMV_THREADS": 512},
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
