"""Pass-1 (symm_panel_kernel) time per tile as a function of the matrix size: how much does the ragged last
super-block row (nb % 8 != 0) cost?  usage: python tools/panel_shape_probe.py [--naux 1184] [--option k=v ...]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from nbed_b200.backend import B200Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--naux", type=int, default=1184)
ap.add_argument("--sizes", default="1280,1376,1408,1536,1024,688")
ap.add_argument("--nocc", type=int, default=5)
ap.add_argument("--option", action="append", default=[])
a = ap.parse_args()
rng = np.random.default_rng(0)
for n in [int(x) for x in a.sizes.split(",")]:
    ctx = B200Context(0)
    for kv in a.option:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.cderi_alloc(n, a.naux)
    ctx.cderi_synth(1, 1e-3, 0)
    orbs = [rng.normal(size=(n, a.nocc)) / np.sqrt(n) for _ in range(2)]
    best = {}
    for rep in range(4):
        ctx.jk_orbitals(orbs)
        t = ctx.timers()
        for k in ("jk_x", "jk_j", "jk_k", "jk_total"):
            best[k] = min(best.get(k, 1e9), t.get(k, 0.0))
    nb = (n + 31) // 32
    tiles = a.naux * nb * (nb + 1) // 2
    print(f"n={n} nb={nb} (nb%8={nb % 8}) naux={a.naux}: pass 1 {best['jk_x']:.3f} ms = {best['jk_x'] * 1e6 * 148 / tiles:.1f} ns per tile per SM "
          f"({8192 * tiles / best['jk_x'] / 1e9:.0f} GB/s), pass2||gram jk_j {best['jk_j']:.3f} jk_k {best['jk_k']:.3f} total {best['jk_total']:.3f}", flush=True)
    ctx.close()
