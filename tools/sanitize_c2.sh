#!/bin/bash
# compute-sanitizer record for the smallest SCF config that exercises every kernel family:
# memcheck + racecheck on __graft_entry__.smoke() (C2 Huzinaga SCF + ao2mo, hand-rolled mbarrier rings included) and on
# one J/K at a size that takes the multi-stage ring path (n = 200), plus (round 2, second record) a 3-cycle SCF at n = 544
# for the single-shot block-product kernel, the guest pass-2 CTAs and programmatic dependent launch.  Bounded by `timeout`.
#   gpurun --timeout 900 -- 'bash tools/sanitize_c2.sh'
# (The first round-2 record is profiles/sanitizer_r02.log.  Later in round 2 the pool closed compute-sanitizer - "runs under it
# have left GPUs needing a reset" - so the kernels added after that record are covered by their unit tools, the parity
# suite and the bit-reproducibility tests instead: profiles/sanitizer_r02b_refused.log.)
set -u
mkdir -p gpurun_out
cat > /tmp/san_jk.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from nbed_b200 import B200Context, synthetic as syn
from nbed_b200.backend import NBD_HUZINAGA
p = syn.make_problem(n=200, naux=24, nocc=5, n_env=4, seed=2, scale=3.0 / np.sqrt(200 * 24))
ctx = B200Context(0)
ctx.load_cderi(p.cderi())
ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
ctx.set_option("sub_min_nao", 128)   # also run the subspace eigensolver kernels
c, e, d, h, info = ctx.huzinaga_scf(4, 1e-9, 1e-7, True)
print("cycles", info["cycles"], "E", info["energy"])
# the single-shot block-product kernel (8-CTA clusters, bulk copies, programmatic dependent launch) needs n >= 512
p = syn.make_problem(n=544, naux=8, nocc=5, n_env=4, seed=3, scale=3.0 / np.sqrt(544 * 8))
ctx.load_cderi(p.cderi())
ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
c, e, d, h, info = ctx.huzinaga_scf(3, 1e-9, 1e-7, True)
print("n=544 cycles", info["cycles"], "E", info["energy"], "block products", ctx.timer_ms("count:sub_applies"))
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool : smoke()" | tee -a gpurun_out/sanitizer_r02b.log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -12 | tee -a gpurun_out/sanitizer_r02b.log
  echo "== compute-sanitizer --tool $tool : n = 200 Huzinaga SCF (ring + subspace kernels)" | tee -a gpurun_out/sanitizer_r02b.log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_jk.py 2>&1 | grep -v "^$" | tail -12 | tee -a gpurun_out/sanitizer_r02b.log
done
