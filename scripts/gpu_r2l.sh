#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "tapered or clone" 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/r2l_new.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2l_pytest.log
timeout 600 python scripts/bench_secondary.py contact50,multi40,sp3d 2>&1 | grep -v "^$" | tee gpurun_out/r2l_secondary.txt
