#!/bin/bash
# round-2 call Z (1 GPU): A/B of the block-product options at C4 inside the SCF loop
cd "$(dirname "$0")/.."
o=gpurun_out
for opt in "sub_pdl=1" "sub_pdl=0" "sub_apply_variant=2" "sub_apply_variant=3"; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --workload C4 --option $opt > $o/r02z_$opt.json 2> $o/r02z_$opt.err
python - "$opt" <<'PY'
import json,sys
f=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02z_{f}.json').read().strip().splitlines()[-1])
print(f,'value', round(d['value'],2), {k:round(v,3) for k,v in d['stages_ms'].items() if k in ('jk_total','eig_sub','iter_total','initial_guess_amortised')}, d['eigensolver']['matrix_block_products_per_step'], d['checksum']['energy_last_step'])
PY
done
