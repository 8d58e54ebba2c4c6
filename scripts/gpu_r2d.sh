#!/bin/bash
# ncu full capture of the lean fast-only kernel + A/B with clocks
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rod_lean_kernel -s 6 -c 1 -f -o gpurun_out/r2d_lean python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
b() { echo "$1"; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['ms_per_step'], d['roofline']['frac'], d['value'], d['clocks'])"; }
{
b "default" X=1
b "default again" X=1
} 2>&1 | tee gpurun_out/r2d_ab.txt
timeout 900 python -m pytest tests -m gpu -q -k "randomized_assembly or cfg4" -s 2>&1 | grep -E "passed|failed|Error|worst" | tail -20 | tee gpurun_out/r2d_pytest.log
