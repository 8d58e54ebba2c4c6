#!/bin/bash
# COOMM-env pass: new parity tests + API tests of the new ids + headline bench line (must be unchanged)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_muscle_envs_gpu.py -x -q -s 2>&1 | tail -40 | tee gpurun_out/r3a_muscle_tests.log
timeout 300 python -m pytest tests/test_env_api_gpu.py -q -k "Crawl or ArmPush or PullWeight" 2>&1 | tail -15 | tee gpurun_out/r3a_api_tests.log
timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r3a_bench.json
cat gpurun_out/r3a_bench.json | cut -c1-400
