"""Basic-block view of an `ncu --page source --csv` dump: python tools/ncu_blocks.py <src.csv> [min_pct]
Consecutive SASS instructions with the same execution count are one block; prints executed warp-instructions, stall
samples and shared-memory wavefronts (actual / ideal) per block."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iW, iWi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = [r for r in rows if len(r) > iWi and r[iE].isdigit()]
seen, uniq = set(), []
for r in data:   # some ncu versions list the kernel twice
    if r[0] in seen:
        break
    seen.add(r[0]); uniq.append(r)
data = uniq
blocks, cur = [], None
for k, r in enumerate(data):
    toks = r[iS].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    ex, samp, wf, wfi = int(r[iE]), int(r[iSamp] or 0), int(r[iW] or 0), int(r[iWi] or 0)
    if cur is None or cur["ex"] != ex:
        cur = dict(start=k, ex=ex, n=0, ops=collections.Counter(), samp=0, wf=0, wfi=0)
        blocks.append(cur)
    cur["n"] += 1; cur["ops"][op.split(".")[0]] += 1; cur["samp"] += samp; cur["wf"] += wf; cur["wfi"] += wfi
tot = sum(b["ex"] * b["n"] for b in blocks); tots = sum(b["samp"] for b in blocks); totw = sum(b["wf"] for b in blocks)
print(f"{len(data)} SASS instructions, {tot} executed warp-instructions, {tots} samples, {totw} shared wavefronts ({sum(b['wfi'] for b in blocks)} ideal)")
for b in blocks:
    w = b["ex"] * b["n"]
    if w > tot * min_pct / 100 or b["samp"] > tots * min_pct / 100 or b["wf"] > totw * min_pct / 100:
        top = " ".join(f"{o}:{c}" for o, c in b["ops"].most_common(6))
        print(f"@{b['start']:5d} n={b['n']:4d} exec={b['ex']:9d} instr%={100*w/tot:5.1f} samp%={100*b['samp']/max(tots,1):5.1f} "
              f"wf%={100*b['wf']/max(totw,1):5.1f} (x{b['wf']/max(b['wfi'],1):.1f} ideal) {top}")
