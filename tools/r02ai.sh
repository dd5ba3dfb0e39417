#!/bin/bash
# round-2 call AI (1 GPU): one-CTA Jacobi eigensolver for n <= 32 - unit test, whole GPU suite, C1 / C2 bench A/B
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 60 build/small_eigh_test 2>&1 | tee $o/r02ai_small_eigh_test.log
timeout 900 python -m pytest tests -x -q -m gpu > $o/r02ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r02ai_pytest.log
for w in C2 C1; do for opt in "small_eigh=1" "small_eigh=0"; do
timeout 200 python bench.py --steps 20 --warmup 3 --no-extras --workload $w --option $opt > $o/r02ai_${w}_$opt.json 2> $o/r02ai_${w}_$opt.err
python - "${w}_$opt" <<'PY'
import json,sys
f=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02ai_{f}.json').read().strip().splitlines()[-1])
print(f,'value', round(d['value'],1), {k:round(v,3) for k,v in d['stages_ms'].items() if v>0}, d['checksum']['energy_last_step'])
PY
done; done
