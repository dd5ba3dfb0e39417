#!/bin/bash
# round-2 call T (2 GPUs): spin-split subspace eigensolver - 1-GPU parity tests, 2-GPU entry-point check, 2-GPU bench A/B
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scf.py tests/test_gpu_fullsize.py -x -q -m gpu > $o/r02t_pytest.log 2>&1; tail -3 $o/r02t_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py > $o/r02t_mgpu.log 2>&1; tail -10 $o/r02t_mgpu.log
for v in "split1:" "split0:--option dist_sub=0"; do
  name=${v%%:*}; opt=${v#*:}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-extras --steps 20 --warmup 3 $opt > $o/r02t_n2_$name.json 2> $o/r02t_n2_$name.err
  python - "$o/r02t_n2_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk %.2f eig %.3f bcast %.3f guess %.3f'%(s['jk_total'],s['eig_sub'],s['eig_bcast'],s['initial_guess_amortised']), d['eigensolver']['matrix_block_products_per_step'], d['eigensolver']['fallbacks_to_cusolver'], d['checksum']['energy_last_step'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
