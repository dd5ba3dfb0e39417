"""GPU parity at the STATED sizes of BASELINE.json configs 3, 4 and 5.

C4 ((H2O)32/def2-TZVP shape, nao = 1376, naux = 4128) and C5 ((H2O)16 shape, nao = 688, naux = 2064, m = 40) are compared
with committed outputs of the unmodified reference loop / the oracle at exactly these sizes (tests/golden/c4_fullsize.npz,
c5_fullsize.npz, made by tests/golden/make_golden_fullsize.py in the dev container: a full-size oracle iteration takes
~25 s of CPU, too long for the GPU suite).  C3 (ethanol/cc-pVTZ shape, nao = 174, naux = 522) and the mu-shift cases
(nao = 174 and 344) run the oracle live.  Tolerances: 1e-8 Ha on energies, 1e-8 on densities, 1e-10 per MO integral
(BASELINE.json north_star)."""
import os
import sys

import numpy as np
import pytest

from nbed_b200 import synthetic as syn
from nbed_b200.backend import NBD_HUZINAGA, NBD_MU_SHIFT
from oracle import nbed_restatement as nr
from oracle import pyscf_restatement as ps

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import fullsize_common as fc  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
E_TOL = 1e-8   # Ha
D_TOL = 1e-8
ERI_TOL = 1e-10


@pytest.fixture(scope="module")
def c4(ctx):
    gold = dict(np.load(os.path.join(GOLD, "c4_fullsize.npz")))
    cfg, p = syn.bench_problem("C4_h2o32_def2tzvp")
    assert (p.n, p.naux) == (int(gold["n"]), int(gold["naux"])) == (1376, 4128) and p.scale == float(gold["scale"])
    ctx.cderi_alloc(p.n, p.naux)
    ctx.cderi_synth(p.seed, p.scale, 0)  # bit-identical to the rows the golden run streamed (tests/test_host_logic.py)
    return gold, cfg, p


def test_c4_full_size_jk_against_the_reference_run(ctx, c4):
    """One J/K at n = 1376, naux = 4128, o = 5 + 5: the first Fock build of the reference loop."""
    gold, cfg, p = c4
    orbs = [np.ascontiguousarray(gold["jk_orbitals"][s]) for s in range(2)]
    vj, vk = ctx.jk_orbitals(orbs)
    ej = fc.check_matrix(vj, gold, "vj", 1e-10, "J at C4")
    ek = fc.check_matrix(vk, gold, "vk", 1e-10, "K at C4")
    print(f"C4 full-size J/K: max sampled |dJ| = {ej:.2e}, |dK| = {ek:.2e}")


@pytest.mark.parametrize("eig_mode", [1, 0])
def test_c4_full_size_huzinaga_cycles_against_the_reference_run(ctx, c4, eig_mode):
    """The first cycles of the C4 Huzinaga SCF: per-cycle energies to 1e-8 Ha, D and Huz to 1e-8 (both eigensolver modes)."""
    gold, cfg, p = c4
    cycles = int(gold["cycles"])
    ctx.set_option("eig_mode", eig_mode)
    try:
        ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
        c, e, d, h, info = ctx.huzinaga_scf(cycles, 1e-14, 1e-14, True)
    finally:
        ctx.set_option("eig_mode", 1)
    assert info["cycles"] == cycles and not info["converged"]
    de = np.abs(info["trace"][:, :2] - gold["energies"]).max()
    dn = np.abs(info["trace"][:, 2] - gold["norm_dm_diff"]).max()
    assert de < E_TOL, (info["trace"], gold["energies"])
    assert dn < D_TOL
    ed = fc.check_matrix(d, gold, "dm", D_TOL, "D after the last cycle")
    eh = fc.check_matrix(h, gold, "huz", 1e-7, "Huzinaga operator of the last cycle")
    k = gold["mo_energy_occ"].shape[1]
    assert np.abs(e[:, :k] - gold["mo_energy_occ"]).max() < 1e-8
    print(f"C4 full-size SCF (eig_mode {eig_mode}): max |dE| = {de:.2e}, |dD| = {ed:.2e}, |dHuz| = {eh:.2e}")


def test_c5_full_size_ao2mo_against_the_oracle(ctx):
    """(H2O)16 shape, naux = 2064, m = 40: all four (m, m, m, m) blocks, 1e-10 per element."""
    gold = dict(np.load(os.path.join(GOLD, "c5_fullsize.npz")))
    cfg, p = syn.bench_problem("C5_h2o16_def2tzvp")
    m = cfg["m"]
    assert (p.n, p.naux, m) == (int(gold["n"]), int(gold["naux"]), int(gold["m"])) == (688, 2064, 40)
    mos = syn.random_orthonormal_mos(p.ovlp, m, 0)
    ctx.cderi_alloc(p.n, p.naux)
    ctx.cderi_synth(p.seed, p.scale, 0)
    got = ctx.ao2mo(mos[0], mos[1])
    err = fc.check_tensor(got, gold, "two", ERI_TOL, "C5 MO integrals")
    h3 = np.array([p.hcore + p.v_emb[0], p.hcore + p.v_emb[1]])
    one = ctx.one_body(h3, mos[0], mos[1])
    assert np.abs(one - gold["one_body"]).max() < ERI_TOL * max(1.0, np.abs(gold["one_body"]).max())
    print(f"C5 full-size ao2mo: max sampled |d(pq|rs)| = {err:.2e}")


def test_c3_stated_size_huzinaga_iterates(ctx):
    """Ethanol/cc-pVTZ shape at its stated naux = 522 (the iterate tests of test_gpu_scf.py use 120 rows)."""
    cfg = dict(syn.CONFIGS["C3_ethanol_ccpvtz"])
    p = syn.make_problem(seed=0, scale=4.0 / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
    b = p.cderi()
    assert b.shape == (522, 174 * 175 // 2)
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    c1, e1, d1, h1, info = ctx.huzinaga_scf(30, 1e-8, 1e-6, True)
    assert info["converged"] == conv0 and abs(info["cycles"] - len(tr)) <= 1
    for k, t in enumerate(tr[: info["cycles"]]):
        assert np.abs(info["trace"][k, :2] - t["energy"]).max() < E_TOL, k
    assert np.abs(d1 - d0).max() < D_TOL and np.abs(e1 - e0).max() < 1e-8 and np.abs(h1 - h0).max() < 1e-7
    # one J/K at the stated size, element-wise
    orbs = [c0[s][:, : p.nocc] for s in range(2)]
    vj, vk = ctx.jk_orbitals(orbs)
    rj, rk = ps.df_get_jk_occ(b, orbs)
    assert np.abs(vj - rj).max() < 1e-11 and np.abs(vk - rk).max() < 1e-11


@pytest.mark.parametrize("n,naux,nocc,n_env", [(174, 522, 9, 4), (344, 1032, 5, 40)])
@pytest.mark.parametrize("mu", [1e6])
def test_mu_shift_at_config_sizes(ctx, n, naux, nocc, n_env, mu):
    """mu-shift projector path (driver.py:500-538) at the C3 shape and at half the C5 shape, mu = 1e6, 1e-8 Ha."""
    import scipy.linalg

    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=0, scale=3.0 / np.sqrt(n * naux))
    b = p.cderi()
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, e_nuc=1.25, max_cycle=40, conv_tol=1e-8)
    _, c = scipy.linalg.eigh(p.hcore, p.ovlp)
    dm0 = np.array([c[:, :nocc] @ c[:, :nocc].T] * 2)
    tr = []
    mf, v_emb = nr.mu_embed(mf, p.v_emb, p.dm_enviro, mu_level_shift=mu, dm0=dm0, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_MU_SHIFT, mu)
    c1, e1, occ1, d1, vhf1, info = ctx.mu_scf(40, 1e-8, 1.25, dm0)
    assert info["converged"] == mf.converged and len(info["trace"]) == len(tr)
    worst = max(abs(info["trace"][k, 0] - t[0]) for k, t in enumerate(tr))
    print(f"mu-shift n = {n}: {len(tr)} cycles, max per-cycle |dE| = {worst:.2e}, final |dE| = {abs(info['e_tot'] - mf.e_tot):.2e}")
    assert worst < E_TOL and abs(info["e_tot"] - mf.e_tot) < E_TOL
    dref = np.asarray(mf.make_rdm1())
    assert np.abs(d1 - dref).max() < D_TOL
    assert np.array_equal(occ1, mf.mo_occ)
