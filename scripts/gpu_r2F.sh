#!/bin/bash
mkdir -p gpurun_out
for knob in "X=1" "SOFTROD_STREAMK=0" "SOFTROD_RODSYNC=0" "SOFTROD_LEAN_FILTER=0" "SOFTROD_PACKED_THREADS=256"; do
  echo "== $knob"; env $knob timeout 600 python scripts/diag_fallback.py pend3d 24 2>&1 | grep "^step 23\|peak" | cut -c1-70
done | tee gpurun_out/r2F_diag3d.txt
