#!/bin/bash
# lean kernel experiments: per-rod barriers, CTA shapes, latency probe
set -u
mkdir -p gpurun_out
python -c "
from gym_softrobot_b200 import _native as n
print('latency', n.probe_latency(0))
print('fp64 peak', n.measure_fp64_peak(0), n.measure_fp64_peak(0, True))
" 2>&1 | tee gpurun_out/r2c_latency.txt
b() { echo "$1"; shift; env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['ms_per_step'], d['roofline']['frac'], d['value'])"; }
{
b "default (rodsync=1 streamk=1)" X=1
b "rodsync=0" SOFTROD_RODSYNC=0
b "rodsync=0 streamk=0" SOFTROD_RODSYNC=0 SOFTROD_STREAMK=0
b "rodsync=1 streamk=0" SOFTROD_STREAMK=0
b "nt=160x3 rodsync=1" SOFTROD_LEAN_THREADS=160
b "nt=160x3 rodsync=0" SOFTROD_LEAN_THREADS=160 SOFTROD_RODSYNC=0
b "nt=320x2 rodsync=1" SOFTROD_LEAN_THREADS=320
b "nt=320x2 rodsync=0" SOFTROD_LEAN_THREADS=320 SOFTROD_RODSYNC=0
b "nt=384 rodsync=1" SOFTROD_PACKED_THREADS=384
b "nt=512 rodsync=1" SOFTROD_PACKED_THREADS=512
b "fastpath=0 (safe kernel only)" SOFTROD_FASTPATH=0
} 2>&1 | tee gpurun_out/r2c_ab.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "golden_substeps or golden_episode or batched_vs_oracle or determinism or full_size or generic_rod or randomized_rods or fast_only or fast_pair" 2>&1 | tail -8 | tee gpurun_out/r2c_pytest.log
timeout 600 python -m pytest tests -m gpu -q -k "randomized_assembly or cfg4" 2>&1 | grep -E "passed|failed|Error|worst" | tail -12 | tee -a gpurun_out/r2c_pytest.log
