#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_omega.py 2>&1 | tail -32 | tee gpurun_out/r2S_diag.txt
