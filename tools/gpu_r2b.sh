set -x
python -m pytest tests/test_gpu_parity2.py -m gpu -q -x -k "tf32" > gpurun_out/r2b_pytest.log 2>&1; echo rc_pytest=$?
tail -5 gpurun_out/r2b_pytest.log
python bench.py --layers > gpurun_out/r2b_bench.log 2> gpurun_out/r2b_bench.err; echo rc_bench=$?
cat gpurun_out/r2b_bench.log; tail -70 gpurun_out/r2b_bench.err
