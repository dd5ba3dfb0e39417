"""Are repeated runs bit-identical?  (development aid)"""
import sys, numpy as np
sys.path.insert(0, '.')
from nbed_b200 import synthetic as syn
from nbed_b200.backend import B200Context, NBD_HUZINAGA
ctx = B200Context(0)
cfg = dict(syn.CONFIGS["C4_h2o32_def2tzvp"], naux=64)
p = syn.make_problem(seed=3, scale=6.0 / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
ctx.cderi_alloc(p.n, p.naux); ctx.cderi_synth(p.seed, p.scale, 0)
for overlap, eig_threads, jv in ((1, 0, 0), (1, 0, 1), (1, 1, 0)):
    for mode in (1, 0):
        ctx.set_option("overlap", overlap); ctx.set_option("eig_mode", mode); ctx.set_option("eig_threads", eig_threads)
        ctx.set_option("jpass_variant", jv)
        runs = []
        for rep in range(4):
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
            c, e, d, h, info = ctx.huzinaga_scf(25, 1e-9, 1e-7, True)
            runs.append((info["trace"].copy(), d.copy(), info["cycles"]))
        def cmp(a, b):
            k = min(len(a), len(b))
            diff = np.abs(a[:k] - b[:k]).max(axis=1)
            nz = np.nonzero(diff > 0)[0]
            return (int(nz[0]) if len(nz) else -1), float(diff.max()), (float(diff[nz[0]]) if len(nz) else 0.0)
        res = [cmp(runs[0][0], r[0]) for r in runs[1:]]
        print(f"overlap={overlap} eig_threads={eig_threads} jpass={jv} eig_mode={mode}: cycles {[r[2] for r in runs]} (first differing cycle, max diff, diff there) = {res}", flush=True)
