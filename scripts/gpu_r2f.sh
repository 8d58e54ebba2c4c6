#!/bin/bash
# full GPU test suite with the new defaults + bench (all configs) + ncu of the 512-thread lean kernel
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2f_pytest.log
timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/r2f_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['ms_per_step'], d['step_ms'], d['roofline']['frac'], d['value'], d['e2e'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
for c in 3 4 5; do timeout 400 python bench.py --config $c --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2f_bench_cfg$c.json; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_cfg$c.json')); print($c, d['ms_per_step'], d['roofline']['frac'], d['value'], d['e2e']['value'], d['cpu_baseline'], d.get('fp32_mode'))"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rod_lean_kernel -s 6 -c 1 -f -o gpurun_out/r2f_lean512 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r2f_bench_reference.json; head -c 600 gpurun_out/r2f_bench_reference.json
