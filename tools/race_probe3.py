"""Which kernel of the pair is wrong?  Pass-2 variants (0 = TMA ring, 1 = LDG) next to the stream-K Gram (development aid)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from nbed_b200.backend import B200Context

n = 1376
naux = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = B200Context(0)
ctx.cderi_alloc(n, naux)
ctx.cderi_synth(3, 3.0 / np.sqrt(n * naux), 0)
rng = np.random.default_rng(0)
orbs = [rng.normal(size=(n, 5)) / np.sqrt(n) for _ in range(2)]
ctx.set_option("overlap", 0); ctx.set_option("syrk", 0)
j0, k0 = ctx.jk_orbitals(orbs)
for jv, syrk, overlap, ctas, safe in ((0, 1, 1, 0, 0), (0, 1, 1, 0, 1), (0, 1, 1, 296, 1), (0, 0, 1, 0, 1), (0, 1, 2, 0, 1), (0, 1, 1, 0, 0)):
    ctx.set_option("jpass_variant", jv); ctx.set_option("syrk", syrk); ctx.set_option("overlap", overlap)
    ctx.set_option("syrk_ctas", ctas); ctx.set_option("jpass_safe", safe)
    out = []
    for r in range(reps):
        j, k = ctx.jk_orbitals(orbs)
        out.append("%.1e/%.1e" % (np.abs(j - j0).max(), np.abs(k - k0).max()))
    print(f"jpass_safe={safe} jpass_variant={jv} syrk={syrk} overlap={overlap} syrk_ctas={ctas}: max|dJ|/max|dK| per rep: {' '.join(out)}", flush=True)
