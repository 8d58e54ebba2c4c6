#!/bin/bash
set -u
mkdir -p gpurun_out
out=gpurun_out/r2z_sanitizer_summary.txt
echo "# compute-sanitizer over scripts/sanitize_smoke.py, round-2 kernels (lean kernel and its four variants incl. the split schedule, generic kernel with tapered rods, fallback pairs, fp32, faithful warp kernel)" > $out
timeout 300 python scripts/sanitize_smoke.py 2>&1 | tail -2 >> $out
for t in memcheck racecheck synccheck; do
  echo "## $t" >> $out
  timeout 1500 compute-sanitizer --tool $t python scripts/sanitize_smoke.py 2>&1 | grep -v "^$" | tail -8 >> $out
done
cat $out
