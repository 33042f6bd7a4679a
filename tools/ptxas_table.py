"""Registers / spills / stack / shared memory per kernel from csrc/ptxas.log (nvcc -Xptxas -v).
usage: python tools/ptxas_table.py [regex]"""
import os, re, subprocess, sys
log = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "nori-ray-tracer_b200", "csrc", "obj", "ptxas.log")).read()
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
cur = None
for line in log.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        stack, spill = m.group(1), (m.group(2), m.group(3))
        continue
    m = re.search(r"Used (\d+) registers.*?(?:(\d+) bytes smem)?$", line)
    if m and cur and (pat is None or pat.search(cur)):
        print(f"{cur[:90]:90s} regs {m.group(1):>3s}  stack {stack:>4s}  spill st/ld {spill[0]:>3s}/{spill[1]:>3s}  smem {m.group(2) or 0}")
