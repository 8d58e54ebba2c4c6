#!/bin/bash
# Standard GPU validation + measurement pass (run under gpurun). Outputs land in gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference.json
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_reference.json", "gpurun_out/${TAG}_bench.json"):
    d = json.load(open(f)); print(f, d["value"], d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], d.get("cpu_baseline"), d.get("clocks"))
PY
# launch list (cold-cache, serialised: compare shares) and one full capture of the top kernel
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rod_packed -s 6 -c 1 -f -o gpurun_out/${TAG}_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
