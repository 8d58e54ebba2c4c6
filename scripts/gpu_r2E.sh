#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/bench_envs.py 2>&1 | grep '^{' | tee gpurun_out/r2E_envs.txt
SOFTROD_LEAN_CONTACT=0 SOFTROD_LEAN_MULTI=0 SOFTROD_LEAN_FILTER=0 SOFTROD_LEAN_SPLINE=0 timeout 900 python scripts/bench_envs.py pend3d,arm,flat,softarm 2>&1 | grep '^{' | tee -a gpurun_out/r2E_envs.txt
