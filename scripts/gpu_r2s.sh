#!/bin/bash
set -u
mkdir -p gpurun_out
for c in multi10 contact50; do
  timeout 700 ncu --set full --clock-control none --import-source on -k regex:rod_lean -c 6 -f -o /tmp/r2s_$c python scripts/bench_secondary.py $c > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/r2s_$c.ncu-rep 0.5 >> gpurun_out/r2s_ncu.txt 2>&1
done
# source-level hot spots of the assembly kernel: stall samples per source line
ncu -i /tmp/r2s_multi10.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('/tmp/src.csv')))
# find header
hi = next(i for i, r in enumerate(rows) if 'Source' in r and any('Sampl' in c for c in r))
hdr = rows[hi]
si = hdr.index('Source'); 
samp = [i for i, c in enumerate(hdr) if c.strip() in ('# Samples', 'Sampling Data (All)', 'Samples')] or [i for i, c in enumerate(hdr) if 'Sampl' in c]
out = open('gpurun_out/r2s_multi10_source.txt', 'w')
out.write(str(hdr) + "\n")
data = []
for r in rows[hi+1:]:
    if len(r) <= max(samp + [si]): continue
    try: v = float(r[samp[0]].replace(',', ''))
    except: continue
    data.append((v, r[si][:150], r))
tot = sum(d[0] for d in data)
out.write(f"total samples {tot}\n")
# keep only launches of the long kernel: the file concatenates kernels; just print top lines
for v, src, r in sorted(data, key=lambda t: -t[0])[:70]:
    out.write(f"{v:10.0f} {100*v/max(tot,1):5.1f}%  {src}\n")
PY
wc -l gpurun_out/r2s_ncu.txt gpurun_out/r2s_multi10_source.txt
