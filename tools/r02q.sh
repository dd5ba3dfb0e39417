#!/bin/bash
# round-2 call Q (2 GPUs): x-cache test, 2-GPU bench with extras, 2-GPU entry-point check
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_scf.py -x -q -m gpu -k "overlap_cache or mu" > $o/r02q_pytest.log 2>&1; tail -3 $o/r02q_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $o/r02q_n2.json 2> $o/r02q_n2.err; tail -2 $o/r02q_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_n2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], {k:round(v,3) for k,v in d['stages_ms'].items()}, d['checksum']['energy_last_step'], 'ao2mo', d['ao2mo']['ms'], d['ao2mo']['device_ms'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py > $o/r02q_mgpu.log 2>&1; tail -12 $o/r02q_mgpu.log
