# one --set full capture of a fused kernel: bash tools/ncu_fused.sh enc|dec <tag>   -> gpurun_out/<tag>_{raw,src}.csv
set -x
W=${1:-enc}; TAG=${2:-nf}
python tools/run_fused.py $W 3 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fused_${W}_kernel --launch-skip 2 -c 1 -o /tmp/${TAG} python tools/run_fused.py $W 3 > gpurun_out/${TAG}_ncu.log 2>&1; echo rc_ncu=$?
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
