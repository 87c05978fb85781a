# ncu passes of one round (B200_PROFILING.md): plain run first, then the launch list and one --set full capture of the
# same command.  usage: bash tools/profile_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
set -x
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --layers"
$CMD > gpurun_out/${TAG}_plain.log 2> gpurun_out/${TAG}_plain.err; echo rc_plain=$?
tail -20 gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 57 -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo rc_list=$?
ncu --set full --clock-control none --import-source on --launch-skip 76 -c 19 -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo rc_full=$?
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full.raw.csv 2>/dev/null; echo rc_raw=$?
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv -k regex:fused_dec > gpurun_out/${TAG}_fused.src.csv 2>/dev/null; echo rc_src1=$?
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv -k regex:f16_first_s2_tma > gpurun_out/${TAG}_first.src.csv 2>/dev/null; echo rc_src2=$?
ls -la gpurun_out/${TAG}_*; du -sh /tmp/${TAG}_full.ncu-rep
