#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_omega.py 2>&1 | grep "fast.*after  500" | cut -c1-170 | tee gpurun_out/r2U_diag.txt
timeout 400 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/r2U_bench.json
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2U_pytest.log
timeout 400 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/r2U_bench.json
