#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2u_pytest.log
for v in 1 0; do
SOFTROD_LEAN_FILTER=$v timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2u_secondary.txt
done
SOFTROD_PACKED_THREADS=256 timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2u_secondary.txt
