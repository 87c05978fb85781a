set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo rc_pytest=$?
tail -40 gpurun_out/r2a_pytest.log
python bench.py --steps 5 --warmup 3 --layers > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; echo rc_bench=$?
cat gpurun_out/r2a_bench.log; tail -30 gpurun_out/r2a_bench.err
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_memcheck.out 2>&1; echo rc_memcheck=$?
tail -15 gpurun_out/r2a_memcheck.log; tail -5 gpurun_out/r2a_memcheck.out
