python bench.py --variant base_model/ch_128 --images 16 --steps 2 --warmup 3 --no-cpu-baseline --layers > gpurun_out/r1g_ch128_plain.log 2>&1; echo rc_plain=$?
ncu --set full --clock-control none --launch-skip 80 -c 40 -o /tmp/r1g_ch128 python bench.py --variant base_model/ch_128 --images 16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_ch128_ncu.log 2>&1; echo rc_full=$?
ncu -i /tmp/r1g_ch128.ncu-rep --page raw --csv > gpurun_out/r1g_ch128.raw.csv 2>/dev/null; echo rc_raw=$?
