#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "octo or assembly or split_schedule or tapered_muscle or fast_only or fast_pair" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2L_pytest.log
timeout 900 python scripts/bench_secondary.py multi40,multi10 2>&1 | grep '^{' | cut -c1-300 | tee gpurun_out/r2L_secondary.txt
timeout 500 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r2L_cfg4.json
