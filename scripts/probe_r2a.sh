#!/bin/bash
# Round-2 first GPU call: (1) is real PyElastica / gymnasium obtainable on the GPU box?  (2) wave / CTA-shape sweep
set -u
mkdir -p gpurun_out
{
  echo "== import probe"; python -c "import elastica, gymnasium; print('elastica', elastica.__version__, 'gymnasium', gymnasium.__version__)" 2>&1 | tail -1
  python -c "import coomm" 2>&1 | tail -1
  echo "== pip download (no index expected)"; timeout 20 python -m pip download --no-deps -d /tmp/pe pyelastica==1.0.0 gymnasium==1.0.0 2>&1 | tail -3
  echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "elast|gymn|coomm|numba" ; echo "(end)"
  echo "== filesystem"; find / -xdev \( -iname "*elastica*" -o -iname "gymnasium*" -o -iname "coomm*" \) -not -path "/proc/*" -not -path "*/repo/*" -not -path "/tmp/*" 2>/dev/null | grep -v "$GRAFT_REPO_ROOT" | head -20; echo "(end)"
  echo "== pip list"; python -m pip list 2>/dev/null | grep -i -E "elast|gymn|coomm|numba"; echo "(end)"
  ls baseline/_ref 2>&1 | head -3
} > gpurun_out/r2a_pyelastica_probe.txt 2>&1
cat gpurun_out/r2a_pyelastica_probe.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
for n in 740 1480 2960 4096 4440 8192; do
  echo "envs=$n" ; timeout 200 python bench.py --envs-per-gpu $n --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['value'])"
done 2>&1 | tee gpurun_out/r2a_waves.txt
for nt in 320 384 512; do
  echo "nt=$nt"; SOFTROD_PACKED_THREADS=$nt timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['value'])"
done 2>&1 | tee gpurun_out/r2a_nt.txt
