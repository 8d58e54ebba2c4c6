#!/bin/bash
# lean kernel v1 (stream-K + trimmed math): parity tests + bench A/B
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_parity_gpu.py::test_continuum_snake_long_horizon_is_a_replica_of_the_reference 2>&1 | tail -25 | tee gpurun_out/r2b_pytest.log
for n in 4096 4440 65536; do
  echo "envs=$n streamk=1"; timeout 200 python bench.py --envs-per-gpu $n --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['value'], d['e2e']['value'])"
done 2>&1 | tee gpurun_out/r2b_bench.txt
echo "envs=4096 streamk=0"; SOFTROD_STREAMK=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['value'])" | tee -a gpurun_out/r2b_bench.txt
echo "envs=4096 fastpath=0"; SOFTROD_FASTPATH=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['value'])" | tee -a gpurun_out/r2b_bench.txt
timeout 300 python scripts/parity_report.py 2>&1 | tail -20 | tee gpurun_out/r2b_parity.txt
