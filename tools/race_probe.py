"""Does the J/K of one fixed input depend on scheduling?  (development aid)
Baseline: serial pass 2 / generic-GEMM Gram.  Then every (syrk, overlap) mode a few times; prints max |dJ|, max |dK|
and where the largest K difference sits (tile coordinates)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from nbed_b200.backend import B200Context

n = 1376
naux = int(sys.argv[1]) if len(sys.argv) > 1 else 1032
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = B200Context(0)
ctx.cderi_alloc(n, naux)
ctx.cderi_synth(3, 3.0 / np.sqrt(n * naux), 0)
rng = np.random.default_rng(0)
orbs = [rng.normal(size=(n, 5)) / np.sqrt(n) for _ in range(2)]
ctx.set_option("overlap", 0); ctx.set_option("syrk", 0)
j0, k0 = ctx.jk_orbitals(orbs)
print("baseline |J| %.3e |K| %.3e" % (np.abs(j0).max(), np.abs(k0).max()), flush=True)
for syrk, overlap in ((0, 1), (1, 0), (1, 1), (1, 2)):
    ctx.set_option("overlap", overlap); ctx.set_option("syrk", syrk)
    for r in range(reps):
        j, k = ctx.jk_orbitals(orbs)
        dj, dk = np.abs(j - j0), np.abs(k - k0)
        s, a, b = np.unravel_index(np.argmax(dk), dk.shape)
        nbad = int((dk > 1e-13 * np.abs(k0).max()).sum())
        print(f"syrk={syrk} overlap={overlap} rep {r}: max|dJ| {dj.max():.3e} max|dK| {dk.max():.3e} at spin {s} ({a},{b}) tile ({a//128},{b//128}); "
              f"K entries off by > 1e-13 rel: {nbad}", flush=True)
