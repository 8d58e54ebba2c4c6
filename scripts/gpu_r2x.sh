#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "3d or split_schedule or pendulum_3d" 2>&1 | tail -5 | tee gpurun_out/r2x_pytest.log
timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2x_secondary.txt
SOFTROD_PACKED_THREADS=256 timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2x_secondary.txt
