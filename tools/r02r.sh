#!/bin/bash
# round-2 call R: persistent Chebyshev filter kernel - parity tests, then A/B bench
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scf.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_golden.py -x -q -m gpu > $o/r02r_pytest.log 2>&1; tail -3 $o/r02r_pytest.log
for v in "pers1:" "pers0:--option sub_persistent=0" "C5_pers1:--workload C5" "C5_pers0:--workload C5 --option sub_persistent=0"; do
  name=${v%%:*}; opt=${v#*:}
  timeout 400 python bench.py --no-extras --steps 20 --warmup 3 $opt > $o/r02r_$name.json 2> $o/r02r_$name.err
  python - "$o/r02r_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk %.2f eig %.3f guess %.3f'%(s['jk_total'],s['eig_sub'],s['initial_guess_amortised']), d['eigensolver']['matrix_block_products_per_step'], d['eigensolver']['fallbacks_to_cusolver'], d['checksum']['energy_last_step'], d['gpu_launches'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
