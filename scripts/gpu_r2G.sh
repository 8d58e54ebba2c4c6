#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "split_schedule" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/r2G_pytest.log
timeout 600 python scripts/diag_fallback.py pend3d 24 2>&1 | grep "^step 23" | cut -c1-80
timeout 600 python scripts/bench_envs.py pend3d 2>&1 | grep '^{' | tee gpurun_out/r2G_envs.txt
