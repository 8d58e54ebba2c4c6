#!/bin/bash
set -u
mkdir -p gpurun_out
for c in contact50 sp3d multi10 snake; do
  timeout 700 ncu --set full --clock-control none -k regex:rod_packed -c 6 -f -o /tmp/r2o_$c python scripts/bench_secondary.py $c > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/r2o_$c.ncu-rep 0.5 >> gpurun_out/r2o_ncu_secondary.txt 2>&1
done
wc -l gpurun_out/r2o_ncu_secondary.txt
