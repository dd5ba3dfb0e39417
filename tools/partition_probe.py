"""Can a subset of the SMs pull the full HBM bandwidth in pass 2?  (development aid for the spatial split of the pair)
C4 shape (n = 1376, naux = 4128, 5 + 5 occupied), serial pair (overlap = 0): pass 2 restricted to SMs with smid % mod < keep."""
import sys
import numpy as np
sys.path.insert(0, '.')
from nbed_b200.backend import B200Context

n, naux = 1376, int(sys.argv[1]) if len(sys.argv) > 1 else 4128
ctx = B200Context(0)
ctx.cderi_alloc(n, naux); ctx.cderi_synth(1, 0.01, 0)
rng = np.random.default_rng(0)
orbs = [rng.normal(size=(n, 5)) / np.sqrt(n) for _ in range(2)]
ctx.set_option("overlap", 0)
gb = naux * (n // 32 + 1) * (n // 32 + 2) // 2 * 8192 / 1e9
for mod, keep, cps in ((0, 0, 2), (4, 3, 2), (4, 2, 2), (4, 2, 3), (4, 1, 2), (4, 1, 3), (5, 2, 3), (3, 1, 3), (2, 1, 3)):
    ctx.set_option("jpass_sm_mod", mod); ctx.set_option("jpass_sm_keep", keep); ctx.set_option("jpass_ctas_per_sm", cps)
    ctx.jk_orbitals(orbs)
    ts = []
    for _ in range(3):
        ctx.jk_orbitals(orbs)
        ts.append(ctx.timers())
    best = min(t['jk_j'] for t in ts)
    frac = 1.0 if mod == 0 else keep / mod
    print(f"pass 2 on {frac:.2f} of the SMs (smid % {mod} < {keep}), {cps} CTAs/SM: {best:.3f} ms = {gb / best * 1e3:.0f} GB/s  "
          f"(Gram alone {min(t['jk_k'] for t in ts):.3f} ms)", flush=True)
