#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_muscle_envs_gpu.py -q -s 2>&1 | tail -30 | tee gpurun_out/r3c_muscle_tests.log
