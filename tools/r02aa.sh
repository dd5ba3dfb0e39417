#!/bin/bash
# round-2 call AA / AQ (2 GPUs): 2-GPU entry-point check and bench A/B (last use: packed-lower all-reduce on / off)
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_scf.py -x -q -m gpu -k "rhf" > $o/r02aa_pytest_rhf.log 2>&1; tail -2 $o/r02aa_pytest_rhf.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py > $o/r02aa_mgpu.log 2>&1; tail -6 $o/r02aa_mgpu.log
for v in "packed1:" "packed0:--option packed_allreduce=0"; do
  name=${v%%:*}; opt=${v#*:}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-extras --steps 20 --warmup 3 $opt > $o/r02aa_n2_$name.json 2> $o/r02aa_n2_$name.err
  python - "$o/r02aa_n2_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk %.2f allreduce %.3f fock %.3f eig %.3f orth %.3f gather %.3f'%(s['jk_total'],s['allreduce'],s['fock'],s['eig_sub'],s['orth'],s['orth_gather']), d['eigensolver']['matrix_block_products_per_step'], d['eigensolver']['fallbacks_to_cusolver'], d['checksum']['energy_last_step'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
