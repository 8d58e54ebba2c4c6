#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "fp32 or tapered" 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/r2m_fp32.log
timeout 300 python scripts/bench_fp32.py 2>&1 | tail -2 | tee gpurun_out/r2m_fp32_bench.txt
timeout 300 python scripts/parity_report.py 2>&1 | tail -6 | tee gpurun_out/r2m_parity.txt
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2m_pytest.log
