set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "model0_symbols or golden or postfilter or cfg5 or encode_decode_every_config or roundtrip_equals or reference_init or fused_crop or full_size" > gpurun_out/q_pytest.log 2>&1; echo rc_pytest=$?
tail -25 gpurun_out/q_pytest.log
timeout 600 python bench.py --layers --no-configs --no-cpu-baseline > gpurun_out/q_bench.log 2> gpurun_out/q_bench.err; echo rc_bench=$?
python -c "
import json
d=json.loads(open('gpurun_out/q_bench.log').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'])
"
tail -22 gpurun_out/q_bench.err
