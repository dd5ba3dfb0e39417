"""Does the subspace eigensolver pay off below n = 256?  (ethanol/cc-pVTZ shape, n = 174)"""
import sys, numpy as np
sys.path.insert(0, '.')
from nbed_b200 import synthetic as syn
from nbed_b200.backend import B200Context, NBD_HUZINAGA
ctx = B200Context(0)
cfg = dict(syn.CONFIGS["C3_ethanol_ccpvtz"])
p = syn.make_problem(seed=1, scale=8.0 / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
ctx.cderi_alloc(p.n, p.naux); ctx.cderi_synth(p.seed, p.scale, 0)
res = {}
for thr in (256, 128):
    ctx.set_option("sub_min_nao", thr)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    ctx.huzinaga_scf(30, 1e-9, 1e-7, True)
    out = ctx.huzinaga_scf(30, 1e-9, 1e-7, True)
    t = ctx.timers()
    res[thr] = out
    print(thr, "cycles", out[4]["cycles"], "scf_total ms", round(t["scf_total"], 2), {k: round(v, 2) for k, v in t.items() if k in ("eigh", "eig_sub", "jk_total")}, "fallbacks", ctx.timer_ms("count:sub_fallbacks"))
print("max |dD|", np.abs(res[256][2] - res[128][2]).max(), "trace", np.abs(res[256][4]["trace"] - res[128][4]["trace"]).max())
