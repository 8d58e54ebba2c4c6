#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/bench_secondary.py sp3d,contact50,snake,multi10 2>&1 | tail -4 | tee -a gpurun_out/r2v_secondary.txt
SOFTROD_FASTPATH=0 timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2v_secondary.txt
SOFTROD_RODSYNC=0 timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2v_secondary.txt
SOFTROD_PACKED_THREADS=256 timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2v_secondary.txt
SOFTROD_PACKED_THREADS=384 timeout 900 python scripts/bench_secondary.py sp3d 2>&1 | tail -1 | tee -a gpurun_out/r2v_secondary.txt
