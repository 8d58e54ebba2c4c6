#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --import-source on -k regex:rod_lean_kernel -s 4 -c 1 -f -o /tmp/r2y_sp3d python scripts/bench_secondary.py sp3d > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/r2y_sp3d.ncu-rep > gpurun_out/r2y_sp3d_ncu.txt 2>&1
python scripts/ncu_hot.py /tmp/r2y_sp3d.ncu-rep 70 > gpurun_out/r2y_sp3d_hot.txt 2>&1
wc -l gpurun_out/r2y_sp3d_hot.txt
