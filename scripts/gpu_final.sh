#!/bin/bash
# last confirmation of the committed build: full GPU suite, smoke, headline bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/final_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/final_smoke.log
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.json; head -c 400 gpurun_out/final_bench.json
