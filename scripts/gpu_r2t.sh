#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r2t_pytest.log
timeout 900 python scripts/bench_secondary.py multi10,multi40 2>&1 | tail -2 | tee -a gpurun_out/r2t_secondary.txt
SOFTROD_PACKED_THREADS=512 timeout 900 python scripts/bench_secondary.py multi10 2>&1 | tail -1 | tee -a gpurun_out/r2t_secondary.txt
