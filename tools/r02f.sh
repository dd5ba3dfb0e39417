#!/bin/bash
# round-2 call F: stream-K Gram parity + A/B benches (syrk on/off, adaptive filter degree on/off)
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_jk.py -x -q -m gpu > $o/r02f_pytest_jk.log 2>&1; tail -3 $o/r02f_pytest_jk.log
for v in "syrk1:" "syrk0:--option syrk=0" "adapt0:--option sub_adaptive=0" "syrk296:--option syrk_ctas=296"; do
  name=${v%%:*}; opt=${v#*:}
  timeout 300 python bench.py --no-extras --steps 20 --warmup 3 $opt > $o/r02f_$name.json 2> $o/r02f_$name.err
  python - "$o/r02f_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk_x %.2f jk_j %.2f jk_k %.2f jk %.2f eig %.3f guess %.3f'%(s['jk_x'],s['jk_j'],s['jk_k'],s['jk_total'],s['eig_sub'],s['initial_guess_amortised']), d['eigensolver']['matrix_block_products_per_step'], d['checksum']['energy_last_step'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
