"""GPU parity: three-/two-centre Coulomb integrals generated on the device (nbd_int3c2e, nbd_cderi_from_basis) against
the oracle's independent McMurchie-Davidson restatement (oracle/gto_restatement.py), and an end-to-end embedded SCF of
the reference's test molecule in cc-pVDZ (BASELINE config 2's molecule and basis) on the device-generated tensor."""
import os

import numpy as np
import pytest
import scipy.linalg

from nbed_b200 import B200UHF, LocalizedSystem, NbdError, huzinaga_scf
from nbed_b200.backend import NBD_HUZINAGA
from oracle import gto_restatement as g
from oracle import nbed_restatement as nr
from oracle import pyscf_restatement as ps

pytestmark = pytest.mark.gpu

WATER = """3

O   0.0000  0.000  0.115
H   0.0000  0.754  -0.459
H   0.0000  -0.754  -0.459
"""  # /root/reference tests/molecules/water.xyz (the geometry of the reference's own tests)


@pytest.fixture(scope="module")
def water():
    atoms = g.parse_xyz(WATER)
    aux = g.even_tempered_aux(g.CCPVDZ, lmax_aux=3)
    atm, bas, env, nbas_ao = g.conc_env(atoms, g.CCPVDZ, aux)
    ao = g.shells_from_env(atm, bas, env, 0, nbas_ao)
    ax = g.shells_from_env(atm, bas, env, nbas_ao)
    j3c, j2c = g.int3c2e_sph(ao, ax), g.int2c2e_sph(ax)
    s, t, v = g.int1e_sph(ao, atoms)
    return dict(atoms=atoms, atm=atm, bas=bas, env=env, nbas_ao=nbas_ao, j3c=j3c, j2c=j2c, ovlp=s, hcore=t + v,
                cderi=g.cholesky_eri(j3c, j2c), e_nuc=g.energy_nuc(atoms))


def test_device_int3c2e_and_int2c2e_match_the_oracle(ctx, water):
    """water / cc-pVDZ (s, p, d orbital shells, 8-primitive contractions) with an even-tempered s-p-d-f auxiliary set."""
    w = water
    j3c, j2c = ctx.int3c2e(w["atm"], w["bas"], w["env"], w["nbas_ao"])
    nao = 24
    il = np.tril_indices(nao)
    assert j3c.shape == (205, nao * (nao + 1) // 2) and j2c.shape == (205, 205)
    want = w["j3c"][:, il[0], il[1]]
    d3 = np.abs(j3c - want).max() / np.abs(want).max()
    d2 = np.abs(j2c - w["j2c"]).max() / np.abs(w["j2c"]).max()
    print(f"int3c2e rel err {d3:.2e}, int2c2e rel err {d2:.2e}")
    assert d3 < 1e-12 and d2 < 1e-12
    assert np.abs(j2c - j2c.T).max() == 0.0


def test_device_cholesky_eri_and_aux_row_ranges(ctx, water):
    """cderi = L^-1 (P|mu nu) generated and decorated on the device == pyscf.df.incore.cholesky_eri restated; any aux
    row range (what a rank of the aux-sharded run holds) equals the same rows of the full tensor."""
    w = water
    nao, naux = ctx.cderi_from_basis(w["atm"], w["bas"], w["env"], w["nbas_ao"])
    assert (nao, naux) == (24, 205)
    full = ctx.cderi_download(0, naux)
    # the metric has condition number ~1e8: forward substitution on either side agrees to ~cond * eps
    assert np.abs(full - w["cderi"]).max() < 1e-8 * np.abs(w["cderi"]).max()
    # the products (mu nu|la si) = sum_P B B, which is what J/K consume, agree much tighter
    il = np.tril_indices(nao)
    eri_dev, eri_ref = full.T @ full, w["cderi"].T @ w["cderi"]
    assert np.abs(eri_dev - eri_ref).max() < 1e-11
    ctx.cderi_from_basis(w["atm"], w["bas"], w["env"], w["nbas_ao"], global_row0=70, naux_local=61)
    part = ctx.cderi_download(0, 61)
    assert np.array_equal(part, full[70:131])
    with pytest.raises(NbdError):
        ctx.cderi_from_basis(w["atm"], w["bas"], w["env"], w["nbas_ao"], global_row0=200, naux_local=10)


def test_water_ccpvdz_embedded_scf_on_device_generated_integrals(ctx, water):
    """End to end on real integrals: DF-UHF of water / cc-pVDZ (pyscf kernel() semantics), then a Huzinaga HF-in-HF
    embedding with the O 1s core as the frozen environment - GPU on the device-generated tensor against the oracle on
    the oracle's tensor, 1e-8 Ha; the embedded density plus the environment reproduces the full HF density."""
    w = water
    s, h = w["ovlp"], w["hcore"]
    ctx.cderi_from_basis(w["atm"], w["bas"], w["env"], w["nbas_ao"])
    _, c = scipy.linalg.eigh(h, s)
    dm0 = np.array([c[:, :5] @ c[:, :5].T] * 2)
    ref = ps.DFUHF(s, h, w["cderi"], (5, 5), e_nuc=w["e_nuc"], max_cycle=50, conv_tol=1e-10)
    conv0, e0, _, c0, _ = ps.scf_kernel(ref, conv_tol=1e-10, dm0=dm0)
    mf = B200UHF(ctx, s, h, (5, 5), e_nuc=w["e_nuc"], max_cycle=50, conv_tol=1e-10)
    e1 = mf.kernel(dm0=dm0)
    assert conv0 and mf.converged and abs(e1 - e0) < 1e-8
    assert -76.03 < e1 < -76.02  # HF / cc-pVDZ water at this geometry (density-fitted)
    # frozen O 1s environment, HF-in-HF: v_emb = J[g_env] - K[g_env]
    c_env = np.array([c0[0][:, :1], c0[1][:, :1]])
    ls = LocalizedSystem(np.arange(1, 5), np.arange(1), np.array([c0[0][:, 1:5], c0[1][:, 1:5]]), c_env,
                         np.array([c0[0][:, :5], c0[1][:, :5]]))
    v_emb = ref.get_veff(dm=np.asarray(ls.dm_enviro))
    act_ref = ps.DFUHF(s, h, w["cderi"], (4, 4), max_cycle=60, conv_tol=1e-10)
    tr = []
    _, eps0, d0, huz0, cv0 = nr.huzinaga_scf(act_ref, v_emb, np.asarray(ls.dm_enviro), dm_conv_tol=1e-8, trace=tr)
    act = B200UHF(ctx, s, h, (4, 4), max_cycle=60, conv_tol=1e-10)
    _, eps1, d1, huz1, cv1, info = huzinaga_scf(act, v_emb, ls.dm_enviro, dm_conv_tol=1e-8, return_info=True)
    assert cv0 and cv1 and abs(info["cycles"] - len(tr)) <= 1
    k = min(info["cycles"], len(tr))
    assert np.abs(info["trace"][:k, :2] - np.array([t["energy"] for t in tr[:k]])).max() < 1e-8
    assert np.abs(np.asarray(d1) - d0).max() < 1e-8 and np.abs(eps1[:, :4] - eps0[:, :4]).max() < 1e-8
    full_dm = np.asarray(ref.make_rdm1())
    assert np.abs(np.asarray(d1) + np.asarray(ls.dm_enviro) - full_dm).max() < 1e-6


def test_f_orbitals_g_auxiliaries_and_general_contractions(ctx):
    """Orbital f shells, auxiliary g shells and a libcint general contraction (nctr = 2, coefficients [nctr][nprim])."""
    atoms = [("O", np.array([0.1, -0.2, 0.3])), ("H", np.array([1.2, 0.9, -0.4]))]
    exps = [3.1, 0.9, 0.35]
    c1, c2 = [0.3, 0.5, 0.4], [-0.2, 0.1, 0.9]
    seg_ao = {"O": [(1, exps, c1), (1, exps, c2), (3, [0.8], [1.0]), (2, [1.3, 0.5], [0.6, 0.5])], "H": [(0, [0.7], [1.0]), (1, [0.9], [1.0])]}
    aux = {"O": [(4, [1.1], [1.0]), (3, [2.0, 0.7], [0.4, 0.7]), (0, [1.5], [1.0])], "H": [(2, [0.8], [1.0]), (1, [1.9], [1.0])]}
    atm, bas, env, nbas_ao = g.conc_env(atoms, seg_ao, aux)
    ao = g.shells_from_env(atm, bas, env, 0, nbas_ao)
    ax = g.shells_from_env(atm, bas, env, nbas_ao)
    want3, want2 = g.int3c2e_sph(ao, ax), g.int2c2e_sph(ax)
    # the same basis with the two p contractions as ONE libcint shell, nctr = 2
    bas_gc = bas.copy()
    env_gc = list(env)
    pc = len(env_gc)
    env_gc += list(env[bas[0, 6] : bas[0, 6] + 3]) + list(env[bas[1, 6] : bas[1, 6] + 3])
    bas_gc[0, 3], bas_gc[0, 6] = 2, pc
    bas_gc = np.delete(bas_gc, 1, axis=0)
    j3c, j2c = ctx.int3c2e(atm, bas_gc, np.array(env_gc), nbas_ao - 1)
    nao = g.nao_sph(ao)
    il = np.tril_indices(nao)
    want = want3[:, il[0], il[1]]
    assert j3c.shape == want.shape == (g.nao_sph(ax), nao * (nao + 1) // 2)
    assert np.abs(j3c - want).max() < 1e-12 * np.abs(want).max() and np.abs(j2c - want2).max() < 1e-12 * np.abs(want2).max()
    # envelope is stated, not silently exceeded: an orbital g shell is refused
    bad = bas.copy()
    bad[2, 1] = 4
    with pytest.raises(NbdError) as ei:
        ctx.int3c2e(atm, bad, env, nbas_ao)
    assert ei.value.code == -4
