#!/bin/bash
# round-2 call AA (2 GPUs): 2-GPU entry-point check and bench after the block-product / read-back changes
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py > $o/r02aa_mgpu.log 2>&1; tail -6 $o/r02aa_mgpu.log
for v in "pdl1:" "pdl0:--option sub_pdl=0"; do
  name=${v%%:*}; opt=${v#*:}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-extras --steps 20 --warmup 3 $opt > $o/r02aa_n2_$name.json 2> $o/r02aa_n2_$name.err
  python - "$o/r02aa_n2_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk %.2f eig %.3f bcast %.3f guess %.3f orth %.3f gather %.3f'%(s['jk_total'],s['eig_sub'],s['eig_bcast'],s['initial_guess_amortised'],s['orth'],s['orth_gather']), d['eigensolver']['matrix_block_products_per_step'], d['eigensolver']['fallbacks_to_cusolver'], d['checksum']['energy_last_step'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
