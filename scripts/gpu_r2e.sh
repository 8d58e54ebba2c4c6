#!/bin/bash
# lean kernel v2 (const trim, peeled last substep): parity + A/B + ncu
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "golden_substeps or golden_episode or batched_vs_oracle or determinism or full_size or generic_rod or randomized_rods or fast_only or fast_pair or free_fall" 2>&1 | tail -8 | tee gpurun_out/r2e_pytest.log
timeout 300 python scripts/parity_report.py 2>&1 | head -6 | tee gpurun_out/r2e_parity.txt
b() { echo "$1"; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['ms_per_step'], d['roofline']['frac'], d['value'], d['clocks'])"; }
{
b "default" X=1
b "rodsync=0" SOFTROD_RODSYNC=0
b "streamk=0" SOFTROD_STREAMK=0
b "nt=512" SOFTROD_PACKED_THREADS=512
b "nt=512 rodsync=0" SOFTROD_PACKED_THREADS=512 SOFTROD_RODSYNC=0
b "nt=320x2" SOFTROD_LEAN_THREADS=320
b "fastpath=0" SOFTROD_FASTPATH=0
echo "envs=65536"; timeout 200 python bench.py --envs-per-gpu 65536 --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ', d['ms_per_step'], d['roofline']['frac'], d['value'])"
} 2>&1 | tee gpurun_out/r2e_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rod_lean_kernel -s 6 -c 1 -f -o gpurun_out/r2e_lean python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "randomized_assembly" -s 2>&1 | grep -E "passed|failed|Error|worst" | tail -30 | tee -a gpurun_out/r2e_pytest.log
