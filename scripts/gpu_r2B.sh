#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "spline or soft_arm" 2>&1 | tail -5 | tee gpurun_out/r2B_pytest.log
timeout 600 python scripts/bench_secondary.py softarm 2>&1 | tail -1 | tee -a gpurun_out/r2B_secondary.txt
timeout 700 ncu --set full --clock-control none -k "regex:rod_(lean|packed)" -c 6 -f -o /tmp/r2B_softarm python scripts/bench_secondary.py softarm > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/r2B_softarm.ncu-rep 0.5 > gpurun_out/r2B_ncu.txt 2>&1
