#!/bin/bash
# round-2 call N: full GPU test-suite + default bench (auto pair split) + other shapes
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $o/r02n_pytest.log 2>&1; tail -3 $o/r02n_pytest.log
for v in "C4:" "C4_split0:--option pair_split=0" "C5:--workload C5" "C5_split0:--workload C5 --option pair_split=0" "C4o20:--workload C4o20"; do
  name=${v%%:*}; opt=${v#*:}
  timeout 400 python bench.py --no-extras --steps 20 --warmup 3 $opt > $o/r02n_$name.json 2> $o/r02n_$name.err
  python - "$o/r02n_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk_x %.2f jk_j %.2f jk_k %.2f jk %.2f eig %.3f guess %.3f'%(s['jk_x'],s['jk_j'],s['jk_k'],s['jk_total'],s['eig_sub'],s['initial_guess_amortised']), d['eigensolver']['matrix_block_products_per_step'], d['checksum']['energy_last_step'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
