"""Compile libsoftrod.so with -Xptxas -v and print registers / spills per kernel."""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_softrobot_b200 import build as b
cmd = [b._nvcc()] + b.NVCC_FLAGS + ["-Xptxas", "-v", "-o", b.LIB_PATH] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
if "error" in out:
    print(out); sys.exit(1)
cur = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void sr::", "")
        spill = None
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur and spill is None:
        spill = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        print(f"{cur:55s} regs={m.group(1):>3s} stack={spill[0]:>5s} spill_st={spill[1]:>5s} spill_ld={spill[2]:>5s}")
        cur = None
