#!/usr/bin/env python
"""Top stalled SASS instructions per kernel from `ncu -i X.ncu-rep --page source --csv` (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kern = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        kern.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for k in kern:
    hdr, data = k["hdr"], k["data"]
    iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iS]) for r in data)
    print("=" * 100)
    print(k["name"][:120], "| samples", tot, "| sass", len(data))
    agg = {}
    for r in data:
        for j in stall:
            agg[hdr[j][6:]] = agg.get(hdr[j][6:], 0) + int(r[j])
    print("  stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:8])
    top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:topn]
    for i in sorted(top):
        r = data[i]
        st = sorted(((hdr[j][6:], int(r[j])) for j in stall if int(r[j]) > 0), key=lambda x: -x[1])[:3]
        print(f"  {i:5d} smp {int(r[iS]):6d} exe {int(r[iE]):9d}  {r[1].strip()[:78]:78s} {st}")
