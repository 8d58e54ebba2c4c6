#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "generic_rod or long_slender or split_schedule" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2M_pytest.log
for v in 1 0; do SOFTROD_LEAN_FOLD=$v timeout 600 python scripts/bench_secondary.py contact512 2>&1 | grep '^{' | cut -c1-300 | tee -a gpurun_out/r2M_secondary.txt; done
timeout 500 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/r2M_cfg5.json
