set -x
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/full_pytest.log 2>&1; echo rc_pytest=$?
tail -8 gpurun_out/full_pytest.log
timeout 900 python bench.py --layers > gpurun_out/full_bench.log 2> gpurun_out/full_bench.err; echo rc_bench=$?
tail -3 gpurun_out/full_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/full_bench.log').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'])
for c in d['configs']: print(c['config'], c['workload'][:60], c['value'], c['ms_per_step'], c['roofline']['kernel'], round(c['roofline']['frac'],3))
print(d['range_coder'])
"
