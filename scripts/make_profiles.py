"""Copy the judged summaries of a gpu_round.sh pass from gpurun_out/ (scratch) into profiles/ (tracked)."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
rep = os.path.join(src, f"{tag}_prof.ncu-rep")
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], stdout=subprocess.PIPE, text=True).stdout
open(os.path.join(dst, f"{tag}_ncu_summary.txt"), "w").write(
    f"# ncu --set full --clock-control none --import-source on -k regex:rod_lean_kernel -s 6 -c 1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline\n" + summ)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
def mb(k):
    v = float(d[k]); unit = u[k]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
traffic = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
json.dump({"tag": tag, "kernel": d.get("Kernel Name"), "dram_bytes_per_launch": traffic,
           "dram_bytes_read": mb("dram__bytes_read.sum"), "dram_bytes_write": mb("dram__bytes_write.sum"),
           "source": f"profiles/{tag}_ncu_summary.txt (ncu --set full, one launch, 4096 envs x 400 substeps)"},
          open(os.path.join(dst, "latest_traffic.json"), "w"), indent=1)
# launch list: keep kernel name + duration only
out = []
for r in csv.reader(open(os.path.join(src, f"{tag}_launches.csv"), errors="ignore")):
    if len(r) > 5 and r[-3] == "gpu__time_duration.sum":
        out.append((r[4].split("(")[0], r[-1]))
with open(os.path.join(dst, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 80 python bench.py --steps 5 --warmup 3 --no-cpu-baseline\nkernel,duration_ns\n")
    for k, v in out: f.write(f"{k},{v}\n")
tot = sum(float(v) for _, v in out)
step = sum(float(v) for k, v in out if "rod_lean" in k or "rod_packed" in k or "rod_substeps" in k)
open(os.path.join(dst, f"{tag}_launches.csv"), "a").write(f"# share of the substep kernel in all profiled launch time: {step / tot:.4f} (excluding the dfma peak probe: {step / (tot - sum(float(v) for k, v in out if 'dfma' in k)):.4f})\n")
for f in (f"{tag}_bench.json", f"{tag}_bench_reference.json", f"{tag}_pytest_gpu.log", f"{tag}_smoke.log", f"{tag}_bench_cfg3.json",
          f"{tag}_bench_cfg4.json", f"{tag}_bench_cfg5.json", f"{tag}_secondary.jsonl", f"{tag}_ncu_variants.txt", f"{tag}_smi.txt"):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
print(open(os.path.join(dst, f"{tag}_ncu_summary.txt")).read()[:600]); print(open(os.path.join(dst, "latest_traffic.json")).read()); print(open(os.path.join(dst, f"{tag}_launches.csv")).read()[-400:])
