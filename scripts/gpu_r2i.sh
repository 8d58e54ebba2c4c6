#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== default (rodsync=1)"; timeout 600 python scripts/bench_secondary.py
echo "== rodsync=0"; SOFTROD_RODSYNC=0 timeout 600 python scripts/bench_secondary.py
echo "== nt=512 rodsync=1"; SOFTROD_PACKED_THREADS=512 timeout 600 python scripts/bench_secondary.py contact50,sp3d,snake,softarm
echo "== nt=384 rodsync=1"; SOFTROD_PACKED_THREADS=384 timeout 600 python scripts/bench_secondary.py contact50,sp3d
} 2>&1 | grep -v "^$" | tee gpurun_out/r2i_secondary.txt
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2i_pytest.log
