#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2H_pytest.log
timeout 600 python scripts/diag_fallback.py arm 10 2>&1 | tail -3 | cut -c1-100
timeout 600 python scripts/bench_envs.py arm,flat 2>&1 | grep '^{' | tee gpurun_out/r2H_envs.txt
timeout 600 python scripts/bench_secondary.py contact50,snake,contact512 2>&1 | grep '^{' | tee gpurun_out/r2H_secondary.txt
