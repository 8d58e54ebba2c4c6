"""Per-instruction stall samples of one kernel launch (ncu --page source), top N SASS lines with their address."""
import csv, subprocess, sys, io
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
ia, isrc = hdr.index("Address"), hdr.index("Source")
isamp = hdr.index("# Samples") if "# Samples" in hdr else next(i for i, c in enumerate(hdr) if "Sampl" in c)
data = []
for r in rows[hi + 1:]:
    try: data.append((float(r[isamp].replace(",", "")), r[ia], r[isrc]))
    except Exception: pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
# cumulative by coarse address region (256 instructions) to see which part of the loop costs
order = {a: i for i, (_, a, _) in enumerate(data)}
for v, a, s in sorted(data, key=lambda t: -t[0])[:top]:
    print(f"{v:9.0f} {100 * v / tot:5.2f}%  #{order[a]:5d} {a} {s[:110]}")
print("--- by block of 64 instructions")
for b in range(0, len(data), 64):
    sv = sum(d[0] for d in data[b:b + 64])
    if sv > 0.01 * tot: print(f"  #{b:5d}-{b + 63:5d} {100 * sv / tot:5.1f}%  first: {data[b][2][:70]}")
