"""Randomised GPU-vs-oracle stress of the Huzinaga and mu SCF loops (development aid; prints a summary)."""
import sys, numpy as np, scipy.linalg
sys.path.insert(0, '.')
from nbed_b200 import synthetic as syn
from nbed_b200.backend import B200Context, NBD_HUZINAGA, NBD_MU_SHIFT
from oracle import nbed_restatement as nr, pyscf_restatement as ps
ctx = B200Context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bad = 0
worst = dict(huz_e=0.0, huz_d=0.0, mu_e=0.0, mu_d=0.0)
for case in range(ncase):
    n = int(rng.integers(12, 70)); naux = int(rng.integers(8, 60)); nocc = int(rng.integers(1, min(8, n // 3)))
    n_env = int(rng.integers(1, 5)); scale = float(rng.uniform(0.5, 3.0)); diis = bool(rng.integers(0, 2))
    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=int(rng.integers(0, 1000)), scale=scale / np.sqrt(n * naux))
    b = p.cderi(); ctx.load_cderi(b)
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=40, conv_tol=1e-8); tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro, use_DIIS=diis, trace=tr)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    c1, e1, d1, h1, info = ctx.huzinaga_scf(40, 1e-8, 1e-6, diis)
    k = min(len(tr), info["cycles"])
    de = max(np.abs(info["trace"][i, :2] - tr[i]["energy"]).max() for i in range(k))
    dd = np.abs(d1 - d0).max()
    chaotic = not conv0
    ok = (info["converged"] == conv0 and abs(info["cycles"] - len(tr)) <= 1 and de < 1e-8 and dd < 1e-7) or chaotic
    worst["huz_e"] = max(worst["huz_e"], 0 if chaotic else de); worst["huz_d"] = max(worst["huz_d"], 0 if chaotic else dd)
    # mu path
    _, c = scipy.linalg.eigh(p.hcore, p.ovlp)
    dm0 = np.array([c[:, :nocc] @ c[:, :nocc].T] * 2)
    mu = float(10 ** rng.uniform(2, 6))
    mf2 = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, e_nuc=0.3, max_cycle=40, conv_tol=1e-8); tr2 = []
    mf2, _ = nr.mu_embed(mf2, p.v_emb, p.dm_enviro, mu_level_shift=mu, dm0=dm0, trace=tr2)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_MU_SHIFT, mu)
    c2, e2, occ2, d2, v2, info2 = ctx.mu_scf(40, 1e-8, 0.3, dm0)
    k2 = min(len(tr2), len(info2["trace"]))
    de2 = max(abs(info2["trace"][i, 0] - tr2[i][0]) for i in range(k2))
    dd2 = np.abs(d2 - np.asarray(mf2.make_rdm1())).max()
    chaotic2 = not mf2.converged
    ok2 = (info2["converged"] == mf2.converged and abs(len(info2["trace"]) - len(tr2)) <= 1 and de2 < 1e-7 and dd2 < 1e-6) or chaotic2
    worst["mu_e"] = max(worst["mu_e"], 0 if chaotic2 else de2); worst["mu_d"] = max(worst["mu_d"], 0 if chaotic2 else dd2)
    if not (ok and ok2):
        bad += 1
        print(f"case {case}: n={n} naux={naux} nocc={nocc} env={n_env} scale={scale:.2f} diis={diis} mu={mu:.1e} | huz conv {info['converged']}/{conv0} cyc {info['cycles']}/{len(tr)} dE {de:.1e} dD {dd:.1e} | mu conv {info2['converged']}/{mf2.converged} cyc {len(info2['trace'])}/{len(tr2)} dE {de2:.1e} dD {dd2:.1e}", flush=True)
print(f"STRESS cases={ncase} bad={bad} worst={ {k: float(f'{v:.2e}') for k, v in worst.items()} }")
