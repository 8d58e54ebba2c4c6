#!/bin/bash
mkdir -p gpurun_out
SOFTROD_PACKED_THREADS=256 timeout 600 python scripts/bench_secondary.py multi10 2>&1 | grep '^{' | cut -c1-300 | tee gpurun_out/r2P.txt
SOFTROD_PACKED_THREADS=256 timeout 600 python scripts/bench_secondary.py contact50,snake 2>&1 | grep '^{' | cut -c1-300 | tee -a gpurun_out/r2P.txt
