#!/bin/bash
# full GPU suite on the build with the muscle envs + env-level throughput of the new envs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r3b_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r3b_smoke.log
timeout 400 python scripts/bench_envs.py push,pull,crawl 2>&1 | grep '^{' | tee gpurun_out/r3b_envs.jsonl
