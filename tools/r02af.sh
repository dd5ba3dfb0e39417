#!/bin/bash
# round-2 call AF (N GPUs): bench with extras at N ranks (N from $1), then the reference arm once (rank 0 only)
cd "$(dirname "$0")/.."
o=gpurun_out
N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 > $o/r02af_n$N.json 2> $o/r02af_n$N.err; tail -3 $o/r02af_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02af_n{N}.json').read().strip().splitlines()[-1])
print('N=',N,'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['checksum']['energy_last_step'], 'ao2mo', d['ao2mo']['ms'], d['ao2mo']['device_ms'], 'clocks', d['clocks'])
PY
