#!/bin/bash
# First GPU call for the experimental 16-consumer-warp panel kernel.  Every step is bounded by
# `timeout`: a hang in the new kernel must not take the box down.
#   gpurun --timeout 600 -- 'bash tools/try_panel16.sh'
set -u
mkdir -p gpurun_out
echo "== parity of the experimental kernel (own contexts, panel_warps = 16)"
NBED_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_gpu_jk.py -m gpu -x -q -k experimental_16_warp 2>&1 | tail -5
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "parity failed or timed out (rc=$rc): not timing it"; exit 0; fi
echo "== A/B at the C4 shape (default 8 warps vs 16 warps)"
for o in "" "--option panel_warps=16"; do
  timeout 150 python bench.py --no-extras $o > gpurun_out/p16.json 2> gpurun_out/p16.log
  python - "$o" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/p16.json").read().strip().splitlines()[-1])
print(sys.argv[1] or "default", round(d["value"], 2), "it/s", {k: round(v, 3) for k, v in d["stages_ms"].items() if k in ("jk_x", "jk_j", "jk_k", "jk_total", "iter_total")},
      "roofline.frac", round(d["roofline"]["frac"], 3))
PY
done
echo "== ncu --set full of one 16-warp launch (only after the run above exited 0)"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:symm_panel16 --launch-skip 1 -c 1 \
  -o gpurun_out/ncu_panel16 python bench.py --no-extras --steps 3 --warmup 3 --option panel_warps=16 > gpurun_out/ncu_panel16.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
