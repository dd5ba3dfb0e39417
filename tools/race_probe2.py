"""Fingerprint of the pass-2 error when it co-runs with the stream-K Gram (development aid)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from nbed_b200.backend import B200Context

n = 1376
naux = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = B200Context(0)
ctx.cderi_alloc(n, naux)
ctx.cderi_synth(3, 3.0 / np.sqrt(n * naux), 0)
rng = np.random.default_rng(0)
orbs = [rng.normal(size=(n, 5)) / np.sqrt(n) for _ in range(2)]
ctx.set_option("overlap", 0); ctx.set_option("syrk", 0)
j0, k0 = ctx.jk_orbitals(orbs)
ctx.set_option("overlap", 1); ctx.set_option("syrk", 1)
bad = None
for r in range(8):
    j, k = ctx.jk_orbitals(orbs)
    dj = j[0] - j0[0]
    print(f"rep {r}: max|dJ| {np.abs(dj).max():.3e}  max|dK| {np.abs(k - k0).max():.3e}", flush=True)
    if np.abs(dj).max() > 1e-12 and bad is None:
        bad = dj.copy()
if bad is None:
    print("no error reproduced"); sys.exit(0)
nb = (n + 31) // 32
tiles = []
for I in range(nb):
    for J in range(I + 1):
        blk = bad[32 * I:32 * I + 32, 32 * J:32 * J + 32]
        if np.abs(blk).max() > 1e-12:
            tiles.append((I, J, float(np.abs(blk).max())))
print(f"{len(tiles)} bad lower tiles of {nb * (nb + 1) // 2}; first 20: {tiles[:20]}")
# rho and the B rows on the host
B = ctx.cderi_download(0, naux)  # [naux][n(n+1)/2]
D = sum(o @ o.T for o in orbs)
iu = np.tril_indices(n)
dtri = D[iu] * 2.0
dtri[iu[0] == iu[1]] *= 0.5
rho = B @ dtri
print("rho[:4]", rho[:4])
idx = np.zeros((n, n), dtype=np.int64)
idx[iu] = np.arange(len(iu[0]))
for (I, J, m) in tiles[:6]:
    r0, c0 = 32 * I, 32 * J
    rows = np.arange(r0, min(n, r0 + 32)); cols = np.arange(c0, min(n, c0 + 32))
    rr, cc = np.meshgrid(rows, cols, indexing="ij")
    hi, lo = np.maximum(rr, cc), np.minimum(rr, cc)
    A = (B[:, idx[hi, lo].ravel()] * rho[:, None]).T  # [elements][naux]
    y = bad[np.ix_(rows, cols)].ravel()
    coef, res, *_ = np.linalg.lstsq(A, y, rcond=None)
    fit = A @ coef
    nz = np.nonzero(np.abs(coef) > 1e-6)[0]
    print(f"tile ({I},{J}) max err {m:.3e}: residual of the best fit by same-tile rows {np.abs(y - fit).max():.3e}; "
          f"rows with |coef| > 1e-6: {nz[:40].tolist()} coef {np.round(coef[nz[:12]], 3).tolist()}")
