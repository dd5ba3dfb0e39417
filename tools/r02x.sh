#!/bin/bash
# round-2 call X (1 GPU): e2e wall-clock breakdown (early export on / off), 32-vector block at C3 / C4 / C5
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 300 python tools/e2e_probe.py > $o/r02x_e2e_probe.log 2>&1; tail -4 $o/r02x_e2e_probe.log
timeout 300 python tools/e2e_probe.py --option early_export=0 > $o/r02x_e2e_probe_noearly.log 2>&1; tail -4 $o/r02x_e2e_probe_noearly.log
for w in C3 C4 C5; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --workload $w --option sub_kb=32 > $o/r02x_${w}_kb32.json 2> $o/r02x_${w}_kb32.err; tail -1 $o/r02x_${w}_kb32.err
done
python - <<'PY'
import json
for f in ['r02x_C3_kb32','r02x_C4_kb32','r02x_C5_kb32']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items() if k in ('jk_total','eig_sub','iter_total','initial_guess_amortised')}, d['eigensolver']['matrix_block_products_per_step'], d['eigensolver']['rayleigh_ritz_per_step'], d['checksum']['energy_last_step'])
    except Exception as e: print(f, 'ERR', e)
PY
