#!/bin/bash
set -u
mkdir -p gpurun_out
out=gpurun_out/r4d_sanitizer_muscle.txt
echo "# compute-sanitizer over scripts/sanitize_muscle.py: muscle-layer instantiation (OctoReach-v0, OctoArmTwo-v0) and the transverse-muscle instantiation (OctoCrawl-v0) of the generic kernel" > $out
timeout 120 python scripts/sanitize_muscle.py 2>&1 | tail -2 >> $out
for t in memcheck racecheck; do
  echo "## $t" >> $out
  timeout 400 compute-sanitizer --tool $t python scripts/sanitize_muscle.py 2>&1 | grep -v "^$" | tail -8 >> $out
done
cat $out
