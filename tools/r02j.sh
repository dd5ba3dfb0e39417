#!/bin/bash
# round-2 call J: A/B benches after the shared-memory carveout hint (stream-K Gram on/off, overlap on/off)
cd "$(dirname "$0")/.."
o=gpurun_out
for v in "syrk1:" "syrk0:--option syrk=0" "syrk1_ser:--option overlap=0" "syrk0_ser:--option syrk=0 --option overlap=0" "syrk1_ov2:--option overlap=2" "syrk1_unsafe:--option jpass_safe=0"; do
  name=${v%%:*}; opt=${v#*:}
  timeout 300 python bench.py --no-extras --steps 20 --warmup 3 $opt > $o/r02j_$name.json 2> $o/r02j_$name.err
  python - "$o/r02j_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['stages_ms']
    print(sys.argv[1], 'it/s %.2f'%d['value'], 'jk_x %.2f jk_j %.2f jk_k %.2f jk %.2f eig %.3f guess %.3f'%(s['jk_x'],s['jk_j'],s['jk_k'],s['jk_total'],s['eig_sub'],s['initial_guess_amortised']), d['eigensolver']['matrix_block_products_per_step'], d['checksum']['energy_last_step'])
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
