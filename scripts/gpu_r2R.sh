#!/bin/bash
mkdir -p gpurun_out
for knob in "X=1" "SOFTROD_LEAN_FILTER=0"; do echo "== $knob"; env $knob timeout 300 python scripts/diag_filter.py 63 1 0 2>&1 | tail -8; done | tee gpurun_out/r2R_diag.txt
