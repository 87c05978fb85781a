#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove a Blackwell-native path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP) in the shipped library.
    python tools/sass_opcodes.py [tf_image_compression_b200/libtic.so] > profiles/rNN_sass_opcodes.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "tf_image_compression_b200" / "libtic.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "REDUX", "ATOMS"]
per = collections.OrderedDict()
cur = None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        base = op.split(".")[0]
        if base in WATCH:
            per[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode counts per kernel: `cuobjdump -sass {Path(lib).name}`\n")
print("| kernel | instructions | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
tot = collections.Counter()
for (k, c), n in zip(per.items(), names):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("tic::", "")
    print(f"| `{n}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")
    tot.update(c)
print(f"| **all** | {tot['_total']} | " + " | ".join(str(tot[w]) for w in WATCH) + " |")
