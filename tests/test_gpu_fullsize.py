"""GPU: size-independent properties at BASELINE.json's full sizes (no oracle run fits these sizes in seconds), error
paths, and less common spin / rank combinations."""
import numpy as np
import pytest

from nbed_b200 import NbdError
from nbed_b200 import synthetic as syn
from nbed_b200.backend import NBD_HUZINAGA, NBD_MU_SHIFT
from oracle import nbed_restatement as nr
from oracle import pyscf_restatement as ps

pytestmark = pytest.mark.gpu


def test_c4_jk_is_additive_over_aux_shards_and_chunks(ctx):
    """(H2O)32/def2-TZVP shape (n = 1376): J/K of the whole aux range == sum over two aux shards (what the multi-GPU
    all-reduce does), also with the half-transformed tensor processed in several aux chunks."""
    n, naux = 1376, 96
    rng = np.random.default_rng(0)
    orbs = [rng.normal(size=(n, 5)) / np.sqrt(n), rng.normal(size=(n, 5)) / np.sqrt(n)]
    ctx.cderi_alloc(n, naux)
    ctx.cderi_synth(7, 0.01, 0)
    vj, vk = ctx.jk_orbitals(orbs)
    ctx.set_option("x_budget_mb", 2)  # ~19 aux rows per chunk
    try:
        vj_c, vk_c = ctx.jk_orbitals(orbs)
    finally:
        ctx.set_option("x_budget_mb", 3072)
    assert np.abs(vj - vj_c).max() < 1e-13 and np.abs(vk - vk_c).max() < 1e-13
    acc_j, acc_k = np.zeros_like(vj), np.zeros_like(vk)
    for lo, hi in ((0, 40), (40, 96)):
        ctx.cderi_alloc(n, hi - lo)
        ctx.cderi_synth(7, 0.01, lo)
        pj, pk = ctx.jk_orbitals(orbs)
        acc_j += pj
        acc_k += pk
    assert np.abs(vj - acc_j).max() < 1e-13 and np.abs(vk - acc_k).max() < 1e-13
    # sampled elements against the definition, from the host copy of a few aux rows
    ctx.cderi_alloc(n, 4)
    ctx.cderi_synth(7, 0.01, 0)
    b = syn.synth_cderi_rows(7, n, 0.01, np.arange(4))
    pj, pk = ctx.jk_orbitals(orbs)
    rj, rk = ps.df_get_jk_occ(b, orbs)
    assert np.abs(pj - rj).max() < 1e-13 and np.abs(pk - rk).max() < 1e-13


def test_c4_huzinaga_scf_invariants_and_eigensolver_modes(ctx):
    """Full n = 1376 embedded SCF on a thin aux range: projector annihilation, idempotency, electron count, and
    agreement of the subspace-tracked run with the cuSOLVER-every-cycle run."""
    cfg = dict(syn.CONFIGS["C4_h2o32_def2tzvp"], naux=64)
    p = syn.make_problem(seed=3, scale=6.0 / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
    ctx.cderi_alloc(p.n, p.naux)
    ctx.cderi_synth(p.seed, p.scale, 0)
    res = {}
    for mode in (1, 0):
        ctx.set_option("eig_mode", mode)
        try:
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
            res[mode] = ctx.huzinaga_scf(25, 1e-9, 1e-7, True)
        finally:
            ctx.set_option("eig_mode", 1)
    c, e, d, h, info = res[1]
    assert info["converged"] and info["cycles"] >= 4
    assert info["cycles"] == res[0][4]["cycles"]
    assert np.abs(info["trace"] - res[0][4]["trace"]).max() < 1e-8
    assert np.abs(d - res[0][2]).max() < 1e-8 and np.abs(e - res[0][1]).max() < 1e-8
    s = p.ovlp
    proj = nr.env_projector(s, p.dm_enviro)
    for sp in range(2):
        assert abs(np.einsum("ij,ji->", d[sp], s) - p.nocc) < 1e-9          # tr(D S) = electrons
        assert np.abs(d[sp] @ s @ d[sp] - d[sp]).max() < 1e-9               # idempotent
        assert abs(np.einsum("ij,ji->", d[sp], proj[sp])) < 1e-9            # no environment character
        assert np.abs(c[sp].T @ s @ c[sp] - np.eye(p.n)).max() < 1e-9       # full orthonormal MO set returned
        assert np.all(np.diff(e[sp]) >= -1e-12)


def test_c5_ao2mo_shape_symmetries_and_sampled_parity(ctx):
    cfg = syn.CONFIGS["C5_h2o16_def2tzvp"]
    n, m, naux = cfg["n"], cfg["m"], 48
    p = syn.make_problem(n=n, naux=naux, nocc=5, n_env=10, seed=1, scale=2.0 / np.sqrt(n * naux))
    b = p.cderi()
    mos = syn.random_orthonormal_mos(p.ovlp, m, 2)
    ctx.load_cderi(b)
    got = ctx.ao2mo(mos[0], mos[1])
    assert got.shape == (4, m, m, m, m)
    chem = got.transpose(0, 1, 4, 2, 3)  # [blk][p, q, r, s] = (pq|rs)
    for blk in range(4):
        assert np.abs(chem[blk] - chem[blk].transpose(1, 0, 2, 3)).max() < 1e-12
        assert np.abs(chem[blk] - chem[blk].transpose(0, 1, 3, 2)).max() < 1e-12
    assert np.abs(chem[0] - chem[0].transpose(2, 3, 0, 1)).max() < 1e-12
    assert np.abs(chem[2] - chem[3].transpose(2, 3, 0, 1)).max() < 1e-12
    want = nr.two_body_integrals(b, mos, restricted=False)
    assert np.abs(got - want).max() < 1e-10 * max(1.0, np.abs(want).max())


def test_error_paths_are_loud(ctx):
    with pytest.raises(NbdError) as ei:
        ctx.cderi_alloc(4000, 2)
        ctx.jk_orbitals([np.ones((4000, 1))])
    assert ei.value.code == -4  # NBD_ERR_UNSUPPORTED: beyond the 3072-AO envelope of the panel kernel
    ctx.cderi_alloc(12, 3)
    with pytest.raises(NbdError):
        ctx.huzinaga_scf(5, 1e-8)  # no nbd_scf_setup yet for this tensor
    with pytest.raises(NbdError):
        ctx.set_option("no_such_option", 1)
    with pytest.raises(NbdError):
        ctx.cderi_upload(np.zeros((5, 78)), 0)  # more rows than allocated


def test_open_shell_and_restricted_mu_path(ctx):
    """nelec_alpha != nelec_beta on the Huzinaga path; the rank-2 mu-shift call is refused loudly."""
    cfg = dict(syn.CONFIGS["C2_h2o_ccpvdz"])
    p = syn.make_problem(seed=0, scale=3.0 / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
    b = p.cderi()
    ctx.load_cderi(b)
    nelec = (4, 2)
    mf = ps.DFUHF(p.ovlp, p.hcore, b, nelec, max_cycle=40, conv_tol=1e-8)
    tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro, trace=tr)
    ctx.scf_setup(nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    c1, e1, d1, h1, info = ctx.huzinaga_scf(40, 1e-8, 1e-6, True)
    assert info["converged"] == conv0 and abs(info["cycles"] - len(tr)) <= 1
    assert np.abs(d1 - d0).max() < 1e-8 and np.abs(e1 - e0).max() < 1e-8
    assert abs(np.trace(d1[0] @ p.ovlp) - 4) < 1e-10 and abs(np.trace(d1[1] @ p.ovlp) - 2) < 1e-10
    # the mu path is spin-resolved only in the reference (driver.py:439 indexes dm_enviro[0]): rank-2 is refused
    ctx.scf_setup((4, 4), p.ovlp, p.hcore, p.v_emb[0], 2.0 * p.dm_enviro[0], NBD_MU_SHIFT, 1e4)
    with pytest.raises(NbdError) as ei:
        ctx.mu_scf(10, 1e-8, 0.0, 2.0 * p.dm_enviro[0])
    assert ei.value.code == -4


def test_repeated_runs_are_bit_identical(ctx):
    """Regression test for a stale-cache bug found in round 1 (rho read through the non-coherent path while pass 2
    co-runs with the K Gram): the same SCF, run three times, must give bit-identical iterates in every mode."""
    cfg = dict(syn.CONFIGS["C4_h2o32_def2tzvp"], naux=64)
    p = syn.make_problem(seed=3, scale=6.0 / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
    ctx.cderi_alloc(p.n, p.naux)
    ctx.cderi_synth(p.seed, p.scale, 0)
    for mode in (1, 0):
        ctx.set_option("eig_mode", mode)
        try:
            runs = []
            for _ in range(3):
                ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
                c, e, d, h, info = ctx.huzinaga_scf(25, 1e-9, 1e-7, True)
                runs.append((info["trace"].copy(), d.copy()))
        finally:
            ctx.set_option("eig_mode", 1)
        for tr, d in runs[1:]:
            assert tr.shape == runs[0][0].shape and np.array_equal(tr, runs[0][0]), mode
            assert np.array_equal(d, runs[0][1]), mode
