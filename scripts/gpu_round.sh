#!/bin/bash
# Standard GPU validation + measurement pass (run under gpurun). Outputs land in gpurun_out/<tag>_*;
# scripts/make_profiles.py <tag> copies the judged summaries into profiles/.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference.json
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
for c in 3 4 5; do timeout 500 python bench.py --config $c --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/${TAG}_bench_cfg$c.json; done
python - <<PY
import json
for f in ["gpurun_out/${TAG}_bench_reference.json", "gpurun_out/${TAG}_bench.json"] + ["gpurun_out/${TAG}_bench_cfg%d.json" % c for c in (3, 4, 5)]:
    try:
        d = json.load(open(f)); print(f, d["value"], d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 900 python scripts/bench_secondary.py 2>&1 | grep '^{' | tee gpurun_out/${TAG}_secondary.jsonl
# launch list (cold-cache, serialised: compare shares) and full captures: the headline kernel (kept as .ncu-rep with
# source) and one launch of every other lean variant + the generic spline kernel (summarised here, text only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rod_lean_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
: > gpurun_out/${TAG}_ncu_variants.txt
for c in contact50 snake multi10 multi40 sp3d contact512 softarm; do
  timeout 700 ncu --set full --clock-control none -k "regex:rod_(lean|packed)" -c 8 -f -o /tmp/${TAG}_$c python scripts/bench_secondary.py $c > /dev/null 2>&1
  echo "## $c" >> gpurun_out/${TAG}_ncu_variants.txt
  python scripts/ncu_summary.py /tmp/${TAG}_$c.ncu-rep 0.5 >> gpurun_out/${TAG}_ncu_variants.txt 2>&1
done
ls -la gpurun_out | tail -20
