"""profiles/<tag>_summary_f16x3.md from the outputs of tools/profile_round.sh <tag>:  python tools/profile_summary.py r2e "note" """
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
G = ROOT / "gpurun_out"
rows = [r for r in csv.reader(open(G / f"{tag}_launches.csv", errors="replace")) if len(r) > 5 and r[0].isdigit()]
ids, hits = [], []
for r in rows:
    if r[0] not in ids:
        ids.append(r[0])
        if "fused_enc_kernel" in ",".join(r):
            hits.append(len(ids) - 1)
step = rows[hits[3]:hits[4]]
tot = sum(float(r[-1]) for r in step) / 1e6
d = json.loads(open(G / f"{tag}_plain.log").read().strip().splitlines()[-1])
lay = d["roofline"]["step_ms_sum_of_layers"]
out = [f"# {tag}: one encode+decode step of BASELINE config 2 under ncu", "",
       f"Command (`tools/profile_round.sh {tag}`, one B200 through `gpurun`): `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --layers`,",
       f"first run plain (exit 0: {d['ms_per_step']:.2f} ms per step = {d['value'] / 1e3:.1f} Gpixel/s device-resident, {d['e2e']['value'] / 1e3:.1f} Gpixel/s host to host; sum of the",
       f"per-layer CUDA-event times {lay:.2f} ms), then `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (launch list:",
       f"`{tag}_launches_f16x3.csv`) and `ncu --set full --clock-control none --import-source on --launch-skip {hits[3]} -c {hits[4] - hits[3]}` (raw page:",
       f"`{tag}_ncu_full_f16x3.csv`).  model_0, 64 images of 2048 x 1536, 12 288 patches of 128 x 128 per launch, fp16-pair tensor path.", note, "",
       "## Launch list of the first timed step (per-launch times are serialised and cold-cache)", "",
       "| launch | kernel | ns | share of the step |", "|---|---|---|---|"]
for r in step:
    out.append(f"| {r[0]} | `{r[4].split('(')[0].replace('void ', '')}` | {int(float(r[-1]))} | {float(r[-1]) / 1e6 / tot * 100:.1f} % |")
k_ms = d["roofline"]["kernel_ms_per_launch"]
out += ["", f"Sum {tot:.3f} ms.  CUDA-event time of the dominant kernel ({d['roofline']['kernel']}) inside the un-profiled run: {k_ms:.3f} ms = {100 * k_ms / lay:.1f} % of",
        f"the {lay:.2f} ms sum of layers (ncu launch list: {100 * float(step[0][-1]) / 1e6 / tot:.1f} %): the shares agree.", "",
        f"## `--set full` capture of the same {len(step)} launches", ""]
out.append(subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_step_summary.py"), str(G / f"{tag}_full.raw.csv")], capture_output=True, text=True).stdout)
(ROOT / "profiles" / f"{tag}_summary_f16x3.md").write_text("\n".join(out))
shutil.copy(G / f"{tag}_launches.csv", ROOT / "profiles" / f"{tag}_launches_f16x3.csv")
shutil.copy(G / f"{tag}_full.raw.csv", ROOT / "profiles" / f"{tag}_ncu_full_f16x3.csv")
print("\n".join(out[-26:]))
