"""Per kernel (substring match): every loop with >= 50 DFMA: instruction mix incl. local-memory traffic."""
import re, subprocess, collections, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gym_softrobot_b200/lib/libsoftrod.so")], stdout=subprocess.PIPE, text=True).stdout
for p in re.split(r"\n\s*Function : ", txt)[1:]:
    name = p.split("\n")[0]
    if not re.search(pat, name):
        continue
    ins = []
    for l in p.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(3)))
    print(name, "total", len(ins))
    for k, (a, op) in enumerate(ins):
        pass
    for l in p.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?(BRA[A-Z.]*)(.*?);", l)
        if m:
            a = int(m.group(1), 16)
            mm = re.search(r"0x([0-9a-f]+)", m.group(4))
            if mm and int(mm.group(1), 16) < a:
                t = int(mm.group(1), 16)
                c = collections.Counter(op.split(".")[0] for ad, op in ins if t <= ad <= a)
                if c["DFMA"] + c["FFMA"] >= 50:
                    n = sum(c.values())
                    print(f"   loop {t:#x}-{a:#x}: {n} instr, FP64 {c['DFMA'] + c['DMUL'] + c['DADD'] + c['DSETP']} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DSETP {c['DSETP']}) "
                          f"LDL {c['LDL']} STL {c['STL']} LDS {c['LDS']} STS {c['STS']} BAR {c['BAR']} MUFU {c['MUFU']} LDC {c['LDC'] + c['LDCU']} R2UR {c['R2UR']} BRA {c['BRA']}")
