#!/bin/bash
# round-2 call P: tests of the overlap cache / mu two-thread solve, copy threads, full bench with extras
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_scf.py tests/test_gpu_jk.py tests/test_gpu_golden.py -x -q -m gpu > $o/r02p_pytest.log 2>&1; tail -3 $o/r02p_pytest.log
timeout 300 python tools/copy_check2.py > $o/r02p_copy.log 2>&1; cat $o/r02p_copy.log
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r02p_bench.json 2> $o/r02p_bench.err; tail -2 $o/r02p_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'mu', d.get('mu_shift',{}).get('ms_per_cycle'), d.get('mu_shift',{}).get('stages_ms_per_cycle'), 'ao2mo', d['ao2mo']['ms'], d['ao2mo']['device_ms'], 'build', d['hamiltonian_build']['wall_ms'], d['hamiltonian_build']['device_ms'], 'veff', d['full_system_veff']['wall_ms'])
PY
