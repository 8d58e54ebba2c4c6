#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/diag_fallback.py flat 6 2>&1 | tail -7 | cut -c1-330 | tee gpurun_out/r2I_diagflat.txt
for v in 1 0; do SOFTROD_FASTPATH=$v timeout 600 python scripts/bench_secondary.py contact50,multi10 2>&1 | grep '^{' | cut -c1-330 | tee -a gpurun_out/r2I_secondary.txt; done
