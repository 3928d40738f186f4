#!/bin/bash
# Run bench.py --quick once per tuning variant of the SpMV kernel; prints ms/step and roofline fraction.
cd "$(dirname "$0")/.."
WL=${1:-ice10k}
for so in hicexplorer_b200/variants/libhicbalance_*.so; do
  name=$(basename $so .so | sed 's/libhicbalance_//')
  HICBALANCE_LIB=$PWD/$so python bench.py --quick --workload $WL --steps 60 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-12s ms/step %.4f  spmv_ms %.4f  GB/s %.0f  frac %.3f  clk %s %s' % ('$name', d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done
