"""GPU parity against the committed golden fixtures (outputs of the unmodified reference + the reference's own
golden energies); needs neither /root/reference nor the oracle's SCF code."""
import os

import numpy as np
import pytest
import scipy.linalg

from nbed_b200 import B200RHF, B200UHF, HamiltonianBuilder, huzinaga_scf, mu_embed
from nbed_b200 import synthetic as syn
from oracle import fock_space as fs
from oracle import pyscf_restatement as ps

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCF_CASES = [("c1", "C1_h2o_sto3g", 2.0), ("c2", "C2_h2o_ccpvdz", 3.0)]


@pytest.fixture(scope="module")
def runs():
    return np.load(os.path.join(GOLD, "reference_runs.npz"))


def scf_problem(key, scale):
    cfg = dict(syn.CONFIGS[key])
    p = syn.make_problem(seed=0, scale=scale / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
    return p, p.cderi()


@pytest.mark.parametrize("name,key,scale", SCF_CASES)
@pytest.mark.parametrize("diis", [True, False])
def test_huzinaga_scf_matches_reference_run(ctx, runs, name, key, scale, diis):
    p, b = scf_problem(key, scale)
    ctx.load_cderi(b)
    mf = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=40, conv_tol=1e-8)
    c, e, d, h, conv = huzinaga_scf(mf, p.v_emb, p.dm_enviro, use_DIIS=diis)
    tag = f"{name}_uhf_diis{int(diis)}"
    assert conv == bool(runs[f"{tag}_conv"])
    assert np.abs(e - runs[f"{tag}_e"]).max() < 1e-8
    assert np.abs(np.asarray(d) - runs[f"{tag}_dm"]).max() < 1e-8
    assert np.abs(h - runs[f"{tag}_huz"]).max() < 1e-7
    assert np.abs(np.abs(c[:, :, : p.nocc]) - runs[f"{tag}_cabs"][:, :, : p.nocc]).max() < 1e-6
    # the tagged density drives the occupied-orbital K route, exactly like pyscf's make_rdm1 output
    assert np.abs(mf.get_veff(dm=d) - mf.get_veff(dm=np.asarray(d))).max() < 1e-10


@pytest.mark.parametrize("name,key,scale", SCF_CASES)
def test_huzinaga_rhf_and_guess_match_reference_run(ctx, runs, name, key, scale):
    p, b = scf_problem(key, scale)
    ctx.load_cderi(b)
    mf = B200RHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=40, conv_tol=1e-8)
    c, e, d, h, conv = huzinaga_scf(mf, p.v_emb[0], 2.0 * p.dm_enviro[0])
    assert conv == bool(runs[f"{name}_rhf_conv"])
    assert np.abs(e - runs[f"{name}_rhf_e"]).max() < 1e-8 and np.abs(np.asarray(d) - runs[f"{name}_rhf_dm"]).max() < 1e-8
    rng = np.random.default_rng(5)
    r = rng.normal(size=(2, p.n, p.n)) * 1e-3
    dm0 = runs[f"{name}_uhf_diis1_dm"] + r + r.transpose(0, 2, 1)
    mf = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=40, conv_tol=1e-8)
    c, e, d, h, conv = huzinaga_scf(mf, p.v_emb, p.dm_enviro, dm_initial_guess=dm0)
    assert np.abs(np.asarray(d) - runs[f"{name}_uhf_guess_dm"]).max() < 1e-8
    mf.get_hcore = lambda *a: p.hcore + p.v_emb
    ee = mf.energy_elec(runs[f"{name}_uhf_diis1_dm"])
    assert np.abs(np.array(ee) - runs[f"{name}_energy_elec"]).max() < 1e-9


def test_spinorb_matches_reference_run(ctx, runs):
    rng = np.random.default_rng(11)
    rng.normal(size=(2, 9, 9)), rng.normal(size=(2, 9, 9))  # keep the generator in step with make_golden.py
    one = rng.normal(size=(2, 3, 3))
    two = rng.normal(size=(4, 3, 3, 3, 3))
    two[0, 0, 1, 2, 0] = 0.99e-8
    two[2, 2, 1, 0, 1] = -1.01e-8
    one[1, 2, 0] = 5e-9
    h1, h2 = ctx.spinorb_from_spatial(one, two, eq_tol=1e-8, two_body_scale=1.0)
    assert np.array_equal(h1, runs["spinorb_h1"]) and np.array_equal(h2, runs["spinorb_h2"])


def test_water_sto3g_reference_goldens(ctx):
    """UHF and FCI energies of the reference's test molecule through the CUDA path (tests/test_driver.py:56-57,76;
    the JW-spectrum == FCI invariant of tests/test_builder.py:55-120)."""
    w = np.load(os.path.join(GOLD, "water_sto3g.npz"))
    s, h, e_nuc = w["S"], w["hcore"], float(w["e_nuc"])
    b = ps.cholesky_eri_exact(w["eri"])  # exact ERI as full-rank DF vectors
    ctx.load_cderi(b)
    _, c = scipy.linalg.eigh(h, s)
    dm0 = np.array([c[:, :5] @ c[:, :5].T] * 2)
    zero = np.zeros((2, 7, 7))
    mf = B200UHF(ctx, s, h, (5, 5), e_nuc=e_nuc, max_cycle=50, conv_tol=1e-11)
    mf, v_emb = mu_embed(mf, zero, zero, 0.0, dm0=dm0)  # no environment: plain pyscf kernel() semantics
    assert mf.converged and abs(mf.e_tot - float(w["ref_e_uhf"])) < 1e-7
    const, h1, h2 = HamiltonianBuilder(mf, e_nuc).build()
    assert h1.shape == (14, 14) and h2.shape == (14,) * 4
    e_fci = fs.ground_energies(const, h1, h2, k=1)[0]
    assert abs(e_fci - float(w["ref_e_fci"])) < 1e-7
