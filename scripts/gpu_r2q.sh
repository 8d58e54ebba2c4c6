#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r2q_pytest.log
timeout 600 python scripts/bench_secondary.py contact50,contact512,snake 2>&1 | tail -3 | tee -a gpurun_out/r2q_secondary.txt
SOFTROD_PACKED_THREADS=384 timeout 600 python scripts/bench_secondary.py contact50,snake 2>&1 | tail -2 | tee -a gpurun_out/r2q_secondary.txt
SOFTROD_PACKED_THREADS=256 timeout 600 python scripts/bench_secondary.py contact50,snake 2>&1 | tail -2 | tee -a gpurun_out/r2q_secondary.txt
