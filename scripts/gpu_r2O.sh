#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "3d or split_schedule" 2>&1 | tail -4 | cut -c1-300
timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | grep '^{' | cut -c1-300 | tee gpurun_out/r2O_secondary.txt
timeout 700 ncu --set full --clock-control none -k regex:rod_lean_kernel -s 4 -c 1 -f -o /tmp/r2O python scripts/bench_secondary.py sp3d > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/r2O.ncu-rep | grep "time_duration\|wavefronts\|bank_conflicts\|pipe_fp64.avg" | tee gpurun_out/r2O_ncu.txt
