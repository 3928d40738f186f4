#!/bin/bash
# Packed (uint16) value layout: ms per ICE iteration / per KR mat-vec for every tuning variant in hicexplorer_b200/variants/
cd "$(dirname "$0")/.."
WL=${1:-ice10k}
for so in hicexplorer_b200/variants/libhicbalance_*.so; do
  name=$(basename $so .so | sed 's/libhicbalance_//')
  HICBALANCE_LIB=$PWD/$so timeout 200 python tools/loop_overhead.py $WL 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)['$WL']
print('%-12s f64 loop %.4f kr %.4f | packed loop %.4f step-level %.4f spmv-only %.4f kr %.4f' % ('$name', d['loop_ms'], d['kr_ms_per_matvec'], d['packed_loop_ms'], d['packed_step_level_ms'], d['packed_spmv_only_ms'], d['packed_kr_ms_per_matvec']))"
done
