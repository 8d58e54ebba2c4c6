#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool synccheck python scripts/sanitize_smoke.py 2>&1 | grep -v "^$" | head -60 > gpurun_out/r2z_synccheck_detail.txt
SOFTROD_RODSYNC=0 timeout 1500 compute-sanitizer --tool synccheck python scripts/sanitize_smoke.py 2>&1 | grep -v "^$" | tail -5 > gpurun_out/r2z_synccheck_norodsync.txt
cat gpurun_out/r2z_synccheck_detail.txt; cat gpurun_out/r2z_synccheck_norodsync.txt
