#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: kernel, registers, stack, spills (tools/ptxas_table.py [log], default stdin)."""
import re
import subprocess
import sys

txt = open(sys.argv[1], errors="replace").read() if len(sys.argv) > 1 else sys.stdin.read()
cur = None
rows = []
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m and cur:
        stack, ss, sl = map(int, m.groups())
        continue_ = (stack, ss, sl)
        rows.append([cur, None, stack, ss, sl])
        continue
    m = re.search(r"Used (\d+) registers", ln)
    if m and rows and rows[-1][1] is None:
        rows[-1][1] = int(m.group(1))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n)
    print(f"{n:70s} regs {r[1]:4d} stack {r[2]:4d} spill st/ld {r[3]:4d}/{r[4]:4d}")
