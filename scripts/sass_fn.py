"""Print the SASS of one kernel (substring match) between two addresses, with line numbers; optional grep pattern."""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
flt = re.compile(sys.argv[4]) if len(sys.argv) > 4 else None
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gym_softrobot_b200/lib/libsoftrod.so")], stdout=subprocess.PIPE, text=True).stdout
for p in re.split(r"\n\s*Function : ", txt)[1:]:
    if pat not in p.split("\n")[0]:
        continue
    k = 0
    for l in p.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m and lo <= int(m.group(1), 16) <= hi:
            k += 1
            if flt is None or flt.search(m.group(2)):
                print(k, m.group(1), m.group(2))
