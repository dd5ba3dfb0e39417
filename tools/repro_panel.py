"""Times the J/K kernels for one shape under different tuning options (development aid)."""
import sys, numpy as np
sys.path.insert(0, '.')
from nbed_b200.backend import B200Context
n, naux = int(sys.argv[1]), int(sys.argv[2])
noccs = [[int(x) for x in a.split(',')] for a in sys.argv[3].split('/')]
stages = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0]
jvars = [int(x) for x in sys.argv[5].split(',')] if len(sys.argv) > 5 else [0]
gtiles = [int(x) for x in sys.argv[6].split(',')] if len(sys.argv) > 6 else [0]
overlaps = [int(x) for x in sys.argv[7].split(',')] if len(sys.argv) > 7 else [1]
ctx = B200Context(0)
ctx.cderi_alloc(n, naux); ctx.cderi_synth(1, 0.01, 0)
rng = np.random.default_rng(0)
for nocc in noccs:
    orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
    for st, jv, gt, ov in [(a, b, g, o) for a in stages for b in jvars for g in gtiles for o in overlaps]:
        ctx.set_option("panel_stages", st)
        ctx.set_option("jpass_variant", jv)
        ctx.set_option("gemm_tile", gt)
        ctx.set_option("overlap", ov)
        ctx.jk_orbitals(orbs)
        ts = []
        for _ in range(3):
            ctx.jk_orbitals(orbs)
            ts.append(ctx.timers())
        best = min(ts, key=lambda t: t['jk_x'])
        print(f"n={n} naux={naux} nocc={nocc} stages={st} jpass={jv} gemm_tile={gt} overlap={ov}: " + " ".join(f"{k}={best.get(k, 0.0):.3f}" for k in ('jk_x','jk_k','jk_j','jk_total')), flush=True)
