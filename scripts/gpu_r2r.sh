#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r2r_pytest.log
for v in 1 0; do
SOFTROD_LEAN_MULTI=$v timeout 900 python scripts/bench_secondary.py multi10,multi40 2>&1 | tail -2 | tee -a gpurun_out/r2r_secondary.txt
done
SOFTROD_PACKED_THREADS=512 timeout 900 python scripts/bench_secondary.py multi10,multi40 2>&1 | tail -2 | tee -a gpurun_out/r2r_secondary.txt
SOFTROD_PACKED_THREADS=768 timeout 900 python scripts/bench_secondary.py multi40 2>&1 | tail -1 | tee -a gpurun_out/r2r_secondary.txt
