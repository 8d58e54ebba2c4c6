#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r2p_pytest.log
for v in 1 0; do
  SOFTROD_LEAN_CONTACT=$v timeout 600 python scripts/bench_secondary.py contact50,contact512,snake 2>&1 | tail -3 | tee -a gpurun_out/r2p_secondary.txt
done
SOFTROD_PACKED_THREADS=384 timeout 600 python scripts/bench_secondary.py contact50,snake 2>&1 | tail -2 | tee -a gpurun_out/r2p_secondary.txt
