set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --layers > gpurun_out/r1g_plain.log 2>&1; echo rc_plain=$?
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 39 -c 60 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_ncu_launch.log 2>&1; echo rc_list=$?
ncu --set full --clock-control none --import-source on --launch-skip 59 -c 20 -o /tmp/r1g_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_ncu_full.log 2>&1; echo rc_full=$?
ncu -i /tmp/r1g_full.ncu-rep --page raw --csv > gpurun_out/r1g_full.raw.csv 2>/dev/null; echo rc_raw=$?
ncu -i /tmp/r1g_full.ncu-rep --page source --csv -k regex:f16_first_s2_tma > gpurun_out/r1g_first.src.csv 2>/dev/null; echo rc_src1=$?
ncu -i /tmp/r1g_full.ncu-rep --page source --csv -k regex:"pair_kernel<.int.4" > gpurun_out/r1g_dec0.src.csv 2>/dev/null; echo rc_src2=$?
ls -la gpurun_out/r1g_*; du -sh /tmp/r1g_full.ncu-rep
