# ncu passes of one round (B200_PROFILING.md): plain run first, then the launch list and one --set full capture of the
# same command.  usage: bash tools/profile_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
set -x
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --layers"
$CMD > gpurun_out/${TAG}_plain.log 2> gpurun_out/${TAG}_plain.err; echo rc_plain=$?
tail -20 gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo rc_list=$?
# first launch of the 4th encode+decode step (= first timed step): index of the 4th fused_enc_kernel in the launch list
read SKIP COUNT < <(python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv", errors="replace")) if len(r) > 5 and r[0].isdigit()]
ids, hits = [], []
for r in rows:
    if r[0] not in ids:
        ids.append(r[0])
        if "fused_enc_kernel" in ",".join(r): hits.append(len(ids) - 1)
print(hits[3], hits[4] - hits[3]) if len(hits) > 4 else print(0, 20)
PY
)
echo "full capture: launch-skip $SKIP, $COUNT launches (one encode+decode step)"
ncu --set full --clock-control none --import-source on --launch-skip $SKIP -c $COUNT -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo rc_full=$?
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full.raw.csv 2>/dev/null; echo rc_raw=$?
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv -k regex:fused_dec > gpurun_out/${TAG}_fused_dec.src.csv 2>/dev/null; echo rc_src1=$?
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv -k regex:fused_enc > gpurun_out/${TAG}_fused_enc.src.csv 2>/dev/null; echo rc_src2=$?
ls -la gpurun_out/${TAG}_*; du -sh /tmp/${TAG}_full.ncu-rep
