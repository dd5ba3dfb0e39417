#!/bin/bash
# round-2 call W (1 GPU): GPU tests after the read-back / Lanczos / early-export changes, then C4 and C3 bench lines
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $o/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $o/r02w_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $o/r02w_n1.json 2> $o/r02w_n1.err; tail -2 $o/r02w_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --workload C3 > $o/r02w_C3.json 2> $o/r02w_C3.err; tail -2 $o/r02w_C3.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --workload C5 > $o/r02w_C5.json 2> $o/r02w_C5.err; tail -2 $o/r02w_C5.err
python - <<'PY'
import json
for f in ['r02w_n1','r02w_C3','r02w_C5']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e') and round(d['e2e']['value'],2), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['eigensolver']['matrix_block_products_per_step'], d['eigensolver']['rayleigh_ritz_per_step'], d['checksum']['energy_last_step'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
