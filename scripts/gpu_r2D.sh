#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "octo or assembly or split_schedule or tapered_muscle or contact or snake or arm_single or tilted" 2>&1 | tail -6 | tee gpurun_out/r2D_pytest.log
timeout 900 python scripts/bench_secondary.py contact50,snake,multi10,multi40,contact512 2>&1 | tail -5 | tee -a gpurun_out/r2D_secondary.txt
