#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r2k_pytest.log
timeout 600 python scripts/bench_secondary.py contact50,snake 2>&1 | grep -v "^$" | tee gpurun_out/r2k_secondary.txt
