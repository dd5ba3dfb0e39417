"""Randomised stress of the subspace-tracked eigensolver path (n >= 256) against the oracle (development aid)."""
import sys, numpy as np
sys.path.insert(0, '.')
from nbed_b200 import synthetic as syn
from nbed_b200.backend import B200Context, NBD_HUZINAGA
from oracle import nbed_restatement as nr, pyscf_restatement as ps
ctx = B200Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nlo, nhi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (256, 420)  # 512+ reaches the single-shot block product
bad = 0; worst_e = worst_d = 0.0
for case in range(ncase):
    n = int(rng.integers(nlo, nhi)); naux = int(rng.integers(8, 40)); nocc = int(rng.integers(1, 22))
    n_env = int(rng.integers(1, 30)); scale = float(rng.uniform(1.0, 5.0)); diis = bool(rng.integers(0, 4))
    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=int(rng.integers(0, 1000)), scale=scale / np.sqrt(n * naux))
    b = p.cderi(); ctx.load_cderi(b)
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8); tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro, use_DIIS=diis, trace=tr)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    a0, f0 = ctx.timer_ms("count:sub_applies"), ctx.timer_ms("count:sub_fallbacks")
    c1, e1, d1, h1, info = ctx.huzinaga_scf(30, 1e-8, 1e-6, diis)
    used, fell = ctx.timer_ms("count:sub_applies") - a0, ctx.timer_ms("count:sub_fallbacks") - f0
    k = min(len(tr), info["cycles"])
    de = max(np.abs(info["trace"][i, :2] - tr[i]["energy"]).max() for i in range(k))
    dd = np.abs(d1 - d0).max(); dspec = np.abs(e1 - e0).max()
    ok = (info["converged"] == conv0 and abs(info["cycles"] - len(tr)) <= 1 and de < 1e-8 and dd < 1e-7 and dspec < 1e-7) or not conv0
    if conv0:
        worst_e, worst_d = max(worst_e, de), max(worst_d, dd)
    print(f"case {case}: n={n} naux={naux} nocc={nocc} env={n_env} scale={scale:.2f} diis={diis} cyc {info['cycles']}/{len(tr)} conv {conv0} applies {used:.0f} fallbacks {fell:.0f} dE {de:.1e} dD {dd:.1e} deps {dspec:.1e} {'OK' if ok else 'BAD'}", flush=True)
    bad += 0 if ok else 1
print(f"STRESS_SUB cases={ncase} bad={bad} worst_dE={worst_e:.2e} worst_dD={worst_d:.2e} cold_starts={ctx.timer_ms('count:sub_cold_starts'):.0f} "
      f"lanczos={ctx.timer_ms('count:sub_lanczos'):.0f} outer={ctx.timer_ms('count:sub_outer'):.0f} fallbacks={ctx.timer_ms('count:sub_fallbacks'):.0f}")
