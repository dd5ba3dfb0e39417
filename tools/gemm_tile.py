import sys, numpy as np
sys.path.insert(0, '.')
from nbed_b200 import synthetic as syn
from nbed_b200.backend import B200Context, NBD_HUZINAGA
ctx = B200Context(0)
cfg = dict(syn.CONFIGS["C4_h2o32_def2tzvp"], naux=256)
p = syn.make_problem(seed=1, scale=16.0 / np.sqrt(cfg["n"] * 4128), **cfg)
ctx.cderi_alloc(p.n, p.naux); ctx.cderi_synth(p.seed, p.scale, 0)
ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
for tile in (0, 128, 64):
    ctx.set_option("gemm_tile", tile)
    ctx.scf_bench_init()
    acc = {}
    for it in range(6):
        ctx.scf_bench_iteration(it)
        for k, v in ctx.timers().items():
            acc[k] = acc.get(k, 0) + v / 6
    print("tile", tile, {k: round(acc[k], 3) for k in ("orth", "fock", "jk_k", "density", "iter_total")})
