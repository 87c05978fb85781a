set -x
N=${1:-2}
nvidia-smi -L
python -m pytest tests -m gpu -q -x -k "nccl or two_devices" > gpurun_out/m${N}_pytest.log 2>&1; echo rc_pytest=$?
tail -5 gpurun_out/m${N}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m${N}_bench.log 2> gpurun_out/m${N}_bench.err; echo rc_bench=$?
tail -3 gpurun_out/m${N}_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/m${N}_bench.log') if l.startswith('{')][-1])
print('weak', d['value'], d['ms_per_step'], 'strong', json.dumps(d['strong']), 'e2e', d['e2e']['value'], d['e2e']['encode_only']['value'], d['e2e']['decode_only']['value'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/pcie_probe.py > gpurun_out/m${N}_pcie.log 2>&1; echo rc_pcie=$?
tail -12 gpurun_out/m${N}_pcie.log
