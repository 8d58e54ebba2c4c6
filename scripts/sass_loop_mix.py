"""Static instruction mix of the main loop of a kernel in libsoftrod.so (cuobjdump -sass)."""
import re, collections, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1] if len(sys.argv) > 1 else "rod_packed_kernelIdLi2E"
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gym_softrobot_b200/lib/libsoftrod.so")], stdout=subprocess.PIPE, text=True).stdout
for p in re.split(r"\n\s*Function : ", txt)[1:]:
    name = p.split("\n")[0]
    if pat not in name: continue
    ins = []
    for l in p.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(3), (m.group(2) or "").strip(), m.group(4)))
    back = []
    for a, op, pred, rest in ins:
        if op.startswith("BRA"):
            mm = re.search(r"0x([0-9a-f]+)", rest)
            if mm and int(mm.group(1), 16) < a: back.append((a, int(mm.group(1), 16)))
    # the substep loop = the smallest backward-branch span that still holds the bulk of the FP64 work
    def n_dfma(t):
        return sum(1 for i in ins if t[1] <= i[0] <= t[0] and i[1].startswith("DFMA"))
    cands = [t for t in back if n_dfma(t) >= 100]
    a1, a0 = min(cands, key=lambda t: t[0] - t[1]) if cands else max(back, key=lambda t: t[0] - t[1])
    loop = [i for i in ins if a0 <= i[0] <= a1]
    c = collections.Counter(i[1].split(".")[0] for i in loop)
    fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(name, "total", len(ins), "loop", len(loop), "fp64", fp64)
    print("  ", c.most_common(30))
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write("\n".join(f"{hex(a)} {pred} {op} {rest}" for a, op, pred, rest in loop))
