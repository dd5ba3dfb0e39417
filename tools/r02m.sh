#!/bin/bash
# round-2 call M: spatial split of the pair [pass 2 || stream-K Gram]
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_jk.py -x -q -m gpu > $o/r02m_pytest_jk.log 2>&1; tail -3 $o/r02m_pytest_jk.log
for ps in 48 52 56 60 64; do
  name=split$ps
  timeout 300 python bench.py --no-extras --steps 20 --warmup 3 --option pair_split=$ps > $o/r02m_$name.json 2> $o/r02m_$name.err
  python - "$o/r02m_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk_x %.2f jk_j %.2f jk_k %.2f jk %.2f eig %.3f guess %.3f'%(s['jk_x'],s['jk_j'],s['jk_k'],s['jk_total'],s['eig_sub'],s['initial_guess_amortised']), d['eigensolver']['matrix_block_products_per_step'], d['checksum']['energy_last_step'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
