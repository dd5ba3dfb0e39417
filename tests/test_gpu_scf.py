"""GPU parity: the embedded-SCF loops (Huzinaga and mu-shift) against the CPU oracle, iterate by iterate."""
import os
import numpy as np
import pytest

from nbed_b200 import synthetic as syn
from nbed_b200.backend import NBD_HUZINAGA, NBD_MU_SHIFT
from oracle import nbed_restatement as nr
from oracle import pyscf_restatement as ps

pytestmark = pytest.mark.gpu

E_TOL = 1e-8  # Ha, BASELINE.json north_star
# (config, coupling scale with DIIS, without DIIS): tuned so the oracle needs 8-26 cycles (DIIS engaged, no chaos)
CASES = [("C1_h2o_sto3g", 2.0, 1.5), ("C2_h2o_ccpvdz", 3.0, 3.0), ("C3_ethanol_ccpvtz", 4.0, 2.0)]


def _same_stop(info, conv0, tr):
    """Same convergence flag and cycle count as the oracle.  The reference stops on |dE| < conv_tol
    (huzinaga_scf.py:196); when |dE| of the deciding cycle sits within rounding noise of conv_tol the two
    FP64 implementations may stop one cycle apart, which is accepted (the iterates themselves must still
    agree to E_TOL on the common prefix, checked by the caller)."""
    assert info["converged"] == conv0
    assert abs(info["cycles"] - len(tr)) <= 1, (info["cycles"], len(tr))


def _problem(name, scale):
    cfg = dict(syn.CONFIGS[name])
    if name == "C3_ethanol_ccpvtz":
        cfg["naux"] = 120  # keeps the NumPy oracle in seconds; the contraction code path is identical
    p = syn.make_problem(seed=0, scale=scale / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
    return p, p.cderi()


@pytest.mark.parametrize("name,scale_diis,scale_plain", CASES)
@pytest.mark.parametrize("use_diis", [True, False])
def test_huzinaga_uhf_iterates(ctx, name, scale_diis, scale_plain, use_diis):
    p, b = _problem(name, scale_diis if use_diis else scale_plain)
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro, use_DIIS=use_diis, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    c1, e1, d1, h1, info = ctx.huzinaga_scf(30, 1e-8, 1e-6, use_diis)
    _same_stop(info, conv0, tr)
    for k, t in enumerate(tr[: info["cycles"]]):
        assert np.abs(info["trace"][k, :2] - t["energy"]).max() < E_TOL, k
        assert abs(info["trace"][k, 2] - t["norm_dm_diff"]) < 1e-8, k
    assert np.abs(e1 - e0).max() < 1e-8
    assert np.abs(d1 - d0).max() < 1e-8
    assert np.abs(h1 - h0).max() < 1e-7
    # MO coefficients up to a phase per column
    for s in range(2):
        ov = np.abs(np.einsum("mi,mn,ni->i", c0[s], p.ovlp, c1[s]))
        assert np.abs(ov[: p.nocc] - 1).max() < 1e-6
    # projector did its job: no environment character in the embedded density
    proj = nr.env_projector(p.ovlp, p.dm_enviro)
    assert abs(np.einsum("sij,sji->", d1, proj)) < 1e-8


def test_huzinaga_rhf_rank2(ctx):
    p, b = _problem("C2_h2o_ccpvdz", 3.0)
    mf = ps.DFRHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    tr = []
    v, g = p.v_emb[0], 2.0 * p.dm_enviro[0]  # spinless convention: doubled density (occupied/base.py:84-85)
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, v, g, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, v, g, NBD_HUZINAGA)
    c1, e1, d1, h1, info = ctx.huzinaga_scf(30, 1e-8, 1e-6, True)
    _same_stop(info, conv0, tr)
    for k, t in enumerate(tr[: info["cycles"]]):
        assert abs(info["trace"][k, 0] - float(t["energy"])) < E_TOL
    assert d1.shape == (p.n, p.n)
    assert np.abs(d1 - d0).max() < 1e-8 and np.abs(h1 - h0).max() < 1e-7 and np.abs(e1 - e0).max() < 1e-8


def test_huzinaga_initial_guess_density(ctx):
    p, b = _problem("C2_h2o_ccpvdz", 3.0)
    rng = np.random.default_rng(5)
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    _, _, dconv, _, _ = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro)
    r = rng.normal(size=dconv.shape) * 1e-3
    dm0 = dconv + r + r.transpose(0, 2, 1)
    tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro, dm_initial_guess=dm0, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    c1, e1, d1, h1, info = ctx.huzinaga_scf(30, 1e-8, 1e-6, True, dm0=dm0)
    _same_stop(info, conv0, tr)
    for k, t in enumerate(tr[: info["cycles"]]):
        assert np.abs(info["trace"][k, :2] - t["energy"]).max() < E_TOL, k
    assert np.abs(d1 - d0).max() < 1e-8


@pytest.mark.parametrize("name,scale,_plain", CASES[:2])
@pytest.mark.parametrize("mu", [1e6, 1e3])
def test_mu_shift_iterates(ctx, name, scale, _plain, mu):
    p, b = _problem(name, scale)
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, e_nuc=1.25, max_cycle=40, conv_tol=1e-8)
    # initial guess: core-Hamiltonian density of the *unshifted* problem (pyscf's minao guess needs basis data)
    import scipy.linalg

    _, c = scipy.linalg.eigh(p.hcore, p.ovlp)
    dm0 = np.array([c[:, : p.nocc] @ c[:, : p.nocc].T] * 2)
    tr = []
    mf, v_emb = nr.mu_embed(mf, p.v_emb, p.dm_enviro, mu_level_shift=mu, dm0=dm0, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_MU_SHIFT, mu)
    c1, e1, occ1, d1, vhf1, info = ctx.mu_scf(40, 1e-8, 1.25, dm0)
    assert info["converged"] == mf.converged
    assert len(info["trace"]) == len(tr)
    tol = E_TOL if mu < 1e5 else 5e-8  # mu = 1e6 amplifies rounding in tr(mu P D)
    for k, t in enumerate(tr):
        assert abs(info["trace"][k, 0] - t[0]) < tol, (k, info["trace"][k], t)
    assert abs(info["e_tot"] - mf.e_tot) < tol
    dref = np.asarray(mf.make_rdm1())
    assert np.abs(d1 - dref).max() < 1e-7
    assert np.array_equal(occ1, mf.mo_occ)
    assert np.abs(e1[:, : p.nocc] - mf.mo_energy[:, : p.nocc]).max() < 1e-6


@pytest.mark.parametrize("n,naux,nocc,n_env,scale", [(300, 40, 5, 12, 4.0), (330, 24, 9, 20, 4.0), (272, 30, 14, 6, 3.0)])
def test_huzinaga_subspace_eigensolver_matches_full_diagonalisation(ctx, n, naux, nocc, n_env, scale):
    """n >= 256: between the first and the last cycle the occupied block is tracked by Chebyshev-filtered subspace
    iteration instead of a full cuSOLVER solve.  Iterates must match both the oracle (full numpy
    eigh every cycle, like the reference) and the library's own eig_mode = 0."""
    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=0, scale=scale / np.sqrt(n * naux))
    b = p.cderi()
    mf = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro, trace=tr)
    assert len(tr) >= 5  # enough cycles for the subspace path to matter
    ctx.load_cderi(b)
    out = {}
    for mode in (0, 1):
        ctx.set_option("eig_mode", mode)
        try:
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
            applies0 = ctx.timer_ms("count:sub_applies")
            out[mode] = ctx.huzinaga_scf(30, 1e-8, 1e-6, True)
            used = ctx.timer_ms("count:sub_applies") - applies0
        finally:
            ctx.set_option("eig_mode", 1)
        assert (used > 0) == (mode == 1)
        c1, e1, d1, h1, info = out[mode]
        _same_stop(info, conv0, tr)
        for k, t in enumerate(tr[: info["cycles"]]):
            assert np.abs(info["trace"][k, :2] - t["energy"]).max() < E_TOL, (mode, k)
            assert abs(info["trace"][k, 2] - t["norm_dm_diff"]) < 1e-8, (mode, k)
        assert np.abs(d1 - d0).max() < 1e-8 and np.abs(h1 - h0).max() < 1e-7
        assert np.abs(e1 - e0).max() < 1e-8  # the FULL spectrum is returned in both modes
        for s in range(2):
            ov = np.abs(np.einsum("mi,mn,ni->i", c0[s], p.ovlp, c1[s]))
            assert np.abs(ov[: p.nocc] - 1).max() < 1e-6
    assert np.abs(out[0][2] - out[1][2]).max() < 1e-9


def test_sc_object_protocol_drives_the_reference_style_loop(ctx):
    """an SCF object whose only device-backed method is get_jk lets the reference's own loop
    (restated line by line in oracle/nbed_restatement.py; the unmodified file produced tests/golden) run its J/K
    on the GPU: get_veff(dm=tagged density) -> get_jk -> nbd_jk.  Same iterates as the all-CPU oracle object."""
    from nbed_b200 import B200RHF, B200UHF

    p, b = _problem("C2_h2o_ccpvdz", 3.0)
    ctx.load_cderi(b)
    cpu = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    gpu = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=30, conv_tol=1e-8)
    t0, t1 = [], []
    r0 = nr.huzinaga_scf(cpu, p.v_emb, p.dm_enviro, trace=t0)
    launches = ctx.launch_count
    r1 = nr.huzinaga_scf(gpu, p.v_emb, p.dm_enviro, trace=t1)
    assert ctx.launch_count > launches and len(t0) == len(t1) and r0[4] == r1[4]
    for a, c in zip(t0, t1):
        assert np.abs(a["energy"] - c["energy"]).max() < 1e-10
    assert np.abs(np.asarray(r0[2]) - np.asarray(r1[2])).max() < 1e-10
    # untagged density -> dense branch (eigen-factorised on the device); get_j; RHF object
    dm = np.asarray(r1[2])
    assert np.abs(gpu.get_veff(dm=dm) - cpu.get_veff(dm=dm)).max() < 1e-10
    assert np.abs(gpu.get_j(dm=dm) - cpu.get_j(dm=dm)).max() < 1e-10
    rc, rg = ps.DFRHF(p.ovlp, p.hcore, b, p.nelec), B200RHF(ctx, p.ovlp, p.hcore, p.nelec)
    assert np.abs(rg.get_veff(dm=dm[0] * 2) - rc.get_veff(dm=dm[0] * 2)).max() < 1e-10
    e_g, e_c = gpu.energy_tot(dm=r1[2]), cpu.energy_tot(dm=r0[2])
    assert abs(e_g - e_c) < 1e-9


@pytest.mark.parametrize("restricted", [False, True])
def test_huzinaga_virtual_orbital_projector(ctx, restricted):
    """dm_environment_virtual (the PAO virtual projector of nbed/driver.py:566-575; second term of
    get_huzinaga_operator, huzinaga_scf.py:82-88) against the oracle, iterate by iterate."""
    from nbed_b200 import B200RHF, B200UHF, get_huzinaga_operator, huzinaga_scf

    p, b = _problem("C2_h2o_ccpvdz", 3.0)
    w, v = np.linalg.eigh(p.ovlp)
    x = (v / np.sqrt(w)) @ v.T
    _, c = np.linalg.eigh(x @ p.hcore @ x)
    cv = (x @ c)[:, -3:]  # three high-lying S-orthonormal orbitals play the environment's virtuals
    gv = np.array([cv @ cv.T, cv[:, :2] @ cv[:, :2].T])
    ctx.load_cderi(b)
    if restricted:
        args = (p.v_emb[0], 2.0 * p.dm_enviro[0], 2.0 * gv[0])
        cpu = ps.DFRHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
        gpu = B200RHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=30, conv_tol=1e-8)
    else:
        args = (p.v_emb, p.dm_enviro, gv)
        cpu = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
        gpu = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=30, conv_tol=1e-8)
    tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(cpu, args[0], args[1], dm_environment_virtual=args[2], trace=tr)
    c1, e1, d1, h1, conv1, info = huzinaga_scf(gpu, args[0], args[1], dm_environment_virtual=args[2], return_info=True)
    _same_stop(info, conv0, tr)
    for k, t in enumerate(tr[: info["cycles"]]):
        assert np.abs(info["trace"][k, : (1 if restricted else 2)] - np.atleast_1d(t["energy"])).max() < E_TOL, k
    assert np.abs(np.asarray(d1) - np.asarray(d0)).max() < 1e-8 and np.abs(h1 - h0).max() < 1e-7
    # the virtual term really contributes, and the host-side formula agrees with the oracle's
    _, _, _, h_occ_only, _ = huzinaga_scf(gpu, args[0], args[1])
    assert np.abs(h1 - h_occ_only).max() > 1e-3
    f = np.random.default_rng(0).normal(size=np.shape(args[1]))
    gs, gvs = args[1] @ p.ovlp, args[2] @ p.ovlp
    assert np.abs(get_huzinaga_operator(f, gs, gvs) - nr.get_huzinaga_operator(f, gs, gvs)).max() < 1e-13


def test_huzinaga_embed_driver_wrapper(ctx):
    """NbedDriver._huzinaga_embed (nbed/driver.py:540-632): SCF + patched core Hamiltonian + e_tot from one more J/K."""
    from nbed_b200 import B200UHF, huzinaga_embed

    p, b = _problem("C2_h2o_ccpvdz", 3.0)
    ctx.load_cderi(b)
    cpu = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, e_nuc=0.75, max_cycle=30, conv_tol=1e-8)
    cpu, v0 = nr.huzinaga_embed(cpu, p.v_emb, p.dm_enviro)
    gpu = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, e_nuc=0.75, max_cycle=30, conv_tol=1e-8)
    gpu, v1 = huzinaga_embed(gpu, p.v_emb, p.dm_enviro)
    assert gpu.converged == cpu.converged and abs(gpu.e_tot - cpu.e_tot) < E_TOL
    assert np.abs(v1 - v0).max() < 1e-7 and np.abs(gpu.get_hcore() - cpu.get_hcore()).max() < 1e-7
    assert np.array_equal(gpu.mo_occ, cpu.mo_occ) and np.abs(gpu.mo_energy - cpu.mo_energy).max() < 1e-8
    assert gpu.get_hcore().shape == (2, p.n, p.n)


def test_subspace_start_and_bound_options_agree(ctx):
    """The cold start of the tracked block (no library eigensolve for the initial guess) and the Lanczos / ||dF||_F
    spectral bounds change how the occupied eigenvectors are found, not what they are: same iterates as the
    cuSOLVER-seeded, row-sum-bounded solver, and the cold path really is taken by default."""
    n, naux, nocc, n_env = 320, 24, 6, 10
    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=3, scale=3.0 / np.sqrt(n * naux))
    ctx.load_cderi(p.cderi())
    runs = {}
    try:
        for cold, bound in ((1, 1), (0, 1), (0, 0)):
            ctx.set_option("sub_cold", cold)
            ctx.set_option("sub_bound", bound)
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
            c0 = ctx.timer_ms("count:sub_cold_starts")
            f0 = ctx.timer_ms("count:sub_fallbacks")
            runs[(cold, bound)] = ctx.huzinaga_scf(30, 1e-8, 1e-6, True)
            assert (ctx.timer_ms("count:sub_cold_starts") - c0 > 0) == (cold == 1)
            if cold == 1:
                assert ctx.timer_ms("count:sub_fallbacks") == f0  # the cold block converged on its own
    finally:
        ctx.set_option("sub_cold", 1)
        ctx.set_option("sub_bound", 1)
    ref = runs[(0, 0)]
    for key in ((1, 1), (0, 1)):
        r = runs[key]
        assert r[4]["cycles"] == ref[4]["cycles"] and r[4]["converged"] == ref[4]["converged"]
        k = r[4]["cycles"]
        assert np.abs(r[4]["trace"][:k, :2] - ref[4]["trace"][:k, :2]).max() < 1e-9
        assert np.abs(r[2] - ref[2]).max() < 1e-9 and np.abs(r[1] - ref[1]).max() < 1e-9


def test_low_rank_environment_projector_matches_the_dense_product(ctx):
    """nbd_scf_set_env_orbitals: F gamma S as (F V)(V^T S) with gamma = V V^T gives the same iterates as the dense
    product (and as the oracle); a factor that does not reproduce dm_env is refused."""
    from nbed_b200 import B200UHF, LocalizedSystem, NbdError, huzinaga_scf

    p = syn.make_problem(n=120, naux=64, nocc=5, n_env=12, seed=4, scale=3.0 / np.sqrt(120 * 64))
    b = p.cderi()
    mf0 = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-9)
    tr = []
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf0, p.v_emb, p.dm_enviro, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    dense = ctx.huzinaga_scf(30, 1e-9, 1e-6, True)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    ctx.scf_set_env_orbitals(p.c_env)
    low = ctx.huzinaga_scf(30, 1e-9, 1e-6, True)
    assert low[4]["cycles"] == dense[4]["cycles"] and low[4]["converged"] == conv0
    assert np.abs(low[4]["trace"] - dense[4]["trace"]).max() < 1e-10
    assert np.abs(low[2] - dense[2]).max() < 1e-10 and np.abs(low[3] - dense[3]).max() < 1e-10
    assert np.abs(low[2] - d0).max() < 1e-8 and np.abs(low[3] - h0).max() < 1e-7
    with pytest.raises(NbdError) as ei:
        ctx.scf_set_env_orbitals(p.c_env * 1.001)
    assert ei.value.code == -2
    # through the mirrored interface: LocalizedSystem.dm_enviro carries c_enviro
    ls = LocalizedSystem(np.arange(5), np.arange(12), p.c_env[:, :, :0], p.c_env, p.c_env)
    assert np.array_equal(np.asarray(ls.dm_enviro), p.dm_enviro) and ls.dm_enviro.factor is not None
    mf = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=30, conv_tol=1e-9)
    c1, e1, d1, h1, conv1 = huzinaga_scf(mf, p.v_emb, ls.dm_enviro)
    assert conv1 == conv0 and np.abs(np.asarray(d1) - d0).max() < 1e-8
    # rank-2 (RHF) convention: doubled density, factor sqrt(2) c
    mfr = ps.DFRHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-9)
    g2 = 2.0 * p.dm_enviro[0]
    _, _, dr, hr, _ = nr.huzinaga_scf(mfr, p.v_emb[0], g2)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb[0], g2, NBD_HUZINAGA)
    ctx.scf_set_env_orbitals(np.sqrt(2.0) * p.c_env[0])
    lowr = ctx.huzinaga_scf(30, 1e-9, 1e-6, True)
    assert np.abs(lowr[2] - dr).max() < 1e-8 and np.abs(lowr[3] - hr).max() < 1e-7


def test_kernel_method_and_spin_resolved_hcore(ctx):
    """B200UHF.kernel() runs pyscf's kernel() semantics with whatever get_hcore returns, so the reference's mu-shift
    sequence (patch get_hcore, call kernel(): driver.py:529-533) binds unchanged; a (2, n, n) patched hcore is accepted
    by the Huzinaga entry as well (a second embedding on an already embedded object)."""
    from nbed_b200 import B200UHF, huzinaga_scf
    import scipy.linalg

    p, b = _problem("C2_h2o_ccpvdz", 3.0)
    mu = 1e5
    _, c = scipy.linalg.eigh(p.hcore, p.ovlp)
    dm0 = np.array([c[:, : p.nocc] @ c[:, : p.nocc].T] * 2)
    ref = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, e_nuc=0.5, max_cycle=40, conv_tol=1e-9)
    ref, v_ref = nr.mu_embed(ref, p.v_emb, p.dm_enviro, mu_level_shift=mu, dm0=dm0)
    ctx.load_cderi(b)
    mf = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, e_nuc=0.5, max_cycle=40, conv_tol=1e-9)
    v_emb = mu * nr.env_projector(p.ovlp, p.dm_enviro) + p.v_emb      # driver.py:518
    hcore_std = mf.get_hcore
    mf.get_hcore = lambda *args: hcore_std(*args) + v_emb              # driver.py:529
    e_tot = mf.kernel(dm0=dm0)                                         # driver.py:533
    assert mf.converged == ref.converged and abs(e_tot - ref.e_tot) < E_TOL
    assert np.array_equal(mf.mo_occ, ref.mo_occ)
    assert np.abs(mf.make_rdm1() - np.asarray(ref.make_rdm1())).max() < 1e-7
    # default guess (no dm0): same fixed point to the convergence threshold
    mf2 = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, e_nuc=0.5, max_cycle=60, conv_tol=1e-10)
    mf2.get_hcore = lambda *args: p.hcore + v_emb
    assert abs(mf2.kernel() - ref.e_tot) < 1e-6 and mf2.converged
    # Huzinaga SCF on an object whose get_hcore is already spin-resolved
    mf3 = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=30, conv_tol=1e-8)
    shift = 0.1 * p.v_emb
    mf3.get_hcore = lambda *args: p.hcore + shift
    r3 = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    r3.get_hcore = lambda *args: p.hcore + shift
    _, _, d_ref, h_ref, conv_ref = nr.huzinaga_scf(r3, p.v_emb, p.dm_enviro)
    _, _, d3, h3, conv3 = huzinaga_scf(mf3, p.v_emb, p.dm_enviro)
    assert conv3 == conv_ref and np.abs(np.asarray(d3) - d_ref).max() < 1e-8 and np.abs(h3 - h_ref).max() < 1e-7
    with pytest.raises(ValueError):
        ctx.scf_setup(p.nelec, p.ovlp, p.hcore[:-1], p.v_emb, p.dm_enviro, NBD_HUZINAGA)


def test_huzinaga_embed_overwrites_virtuals_like_the_driver(ctx):
    """driver.py:604-619 (PAO branch): with c_loc_virt present the occupied block is kept and the rest is re-sliced -
    mirrored verbatim, leading-axis slicing included."""
    from nbed_b200 import B200UHF, LocalizedSystem, huzinaga_embed

    p, b = _problem("C2_h2o_ccpvdz", 3.0)
    ctx.load_cderi(b)
    ls = LocalizedSystem(np.arange(4), np.arange(1), p.c_env[:, :, :0], p.c_env, p.c_env, c_loc_virt=np.zeros((2, p.n, 6)))
    mf = B200UHF(ctx, p.ovlp, p.hcore, p.nelec, max_cycle=30, conv_tol=1e-8)
    mf, v_emb = huzinaga_embed(mf, p.v_emb, ls.dm_enviro, localized_system=ls)
    ref = ps.DFUHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(ref, p.v_emb, p.dm_enviro)
    occ0 = ref.get_occ(e0, c0)
    occ_any = np.sum(occ0, axis=0)
    want = np.concatenate((c0[..., occ_any > 0], c0[..., occ_any == 0][:6]), axis=2)
    assert mf.mo_coeff.shape == want.shape and mf.mo_occ.shape == occ0[: want.shape[-1]].shape
    assert np.abs(np.abs(mf.mo_coeff[..., : p.nocc]) - np.abs(want[..., : p.nocc])).max() < 1e-6
    assert np.abs(v_emb - (h0 + p.v_emb)).max() < 1e-7 and mf.converged == conv0


@pytest.mark.parametrize("n,nocc", [(320, 6), (333, 14), (544, 5), (530, 12)])
def test_cluster_split_k_block_product_matches_the_single_cta_kernel(ctx, n, nocc):
    """The block-product kernels of the subspace eigensolver against each other - 0: automatic choice (the single-shot
    8-CTA-cluster bulk-copy kernel from n = 512 when its grid is resident at once, else the 4-CTA-cluster ring kernel),
    1: one CTA per 16 rows, 2 / 3: the two cluster kernels forced, and programmatic dependent launch switched off: same
    iterates, no fallback to the library eigensolver, no rejected run.  n = 333 is odd (8-byte copy path, ragged last
    row block and chunk; the single-shot kernel needs an even n) with a 32-vector block; n = 530 has a ragged last row
    block and a contraction tail that is not a multiple of 4."""
    naux, n_env = 24, 10
    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=5, scale=3.0 / np.sqrt(n * naux))
    ctx.load_cderi(p.cderi())
    runs = {}
    try:
        for variant in (0, 1, 3, 4) + ((2,) if n % 2 == 0 else ()):
            ctx.set_option("sub_apply_variant", variant % 4)
            ctx.set_option("sub_pdl", 0 if variant == 4 else 1)
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
            f0, r0, a0 = (ctx.timer_ms(k) for k in ("count:sub_fallbacks", "count:sub_rejects", "count:sub_applies"))
            runs[variant] = ctx.huzinaga_scf(30, 1e-8, 1e-6, True)
            assert ctx.timer_ms("count:sub_fallbacks") == f0 and ctx.timer_ms("count:sub_rejects") == r0
            assert ctx.timer_ms("count:sub_applies") > a0
    finally:
        ctx.set_option("sub_apply_variant", 0)
        ctx.set_option("sub_pdl", 1)
    a = runs[0]
    for variant, b in runs.items():
        assert a[4]["cycles"] == b[4]["cycles"] and a[4]["converged"] == b[4]["converged"], variant
        assert np.abs(a[4]["trace"] - b[4]["trace"]).max() < 1e-9, variant
        assert np.abs(a[2] - b[2]).max() < 1e-9 and np.abs(a[1] - b[1]).max() < 1e-8, variant
    # and against the oracle's full diagonalisation in every cycle
    mf = ps.DFUHF(p.ovlp, p.hcore, p.cderi(), p.nelec, max_cycle=30, conv_tol=1e-8)
    _, e0, d0, _, conv0 = nr.huzinaga_scf(mf, p.v_emb, p.dm_enviro)
    assert a[4]["converged"] == conv0 and np.abs(a[2] - d0).max() < 1e-8


def test_overlap_cache_reuses_and_refreshes_x(ctx):
    """nbd_scf_setup keeps X = S^-1/2 when the uploaded overlap is bit-identical to the previous one (the driver embeds
    the same molecule several times, nbed/driver.py:1138-1231) and rebuilds it as soon as one element differs."""
    pa, ba = _problem("C2_h2o_ccpvdz", 3.0)
    ctx.load_cderi(ba)

    def run(ovlp, cache):
        ctx.set_option("x_cache", cache)
        ctx.scf_setup(pa.nelec, ovlp, pa.hcore, pa.v_emb, pa.dm_enviro, NBD_HUZINAGA)
        ms = ctx.timer_ms("eigh")
        c, e, d, h, info = ctx.huzinaga_scf(30, 1e-8, 1e-6, True)
        return e, d, info["trace"].copy(), ms

    try:
        e0, d0, t0, ms0 = run(pa.ovlp, 0)          # no cache: reference result
        e1, d1, t1, ms1 = run(pa.ovlp, 1)          # fills the cache (x_cache was just switched: rebuilt)
        e2, d2, t2, ms2 = run(pa.ovlp, 1)          # reuses X
        assert ms1 > 0.0 and ms2 <= 0.0, (ms1, ms2)
        assert np.array_equal(d1, d0) and np.array_equal(d2, d0) and np.array_equal(t2, t0) and np.array_equal(e2, e0)
        s2 = pa.ovlp.copy()
        s2[0, 1] += 1e-3
        s2[1, 0] += 1e-3
        e3, d3, t3, ms3 = run(s2, 1)               # one element differs: rebuilt
        assert ms3 > 0.0
        e4, d4, t4, ms4 = run(s2, 0)
        assert np.array_equal(d3, d4) and np.array_equal(e3, e4)
        assert np.abs(d3 - d0).max() > 1e-6
        e5, d5, t5, ms5 = run(pa.ovlp, 1)          # back to the first overlap: rebuilt again, same result as before
        assert ms5 > 0.0 and np.array_equal(d5, d0)
    finally:
        ctx.set_option("x_cache", 1)


def _run_cuda_tool(tmp_path, name, args=()):
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / name
    subprocess.run([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe),
                    os.path.join(root, "tools", name + ".cu")], check=True, timeout=600)
    return subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=120)


def test_small_eigh_kernel_against_its_definition(tmp_path):
    """small_eigh_kernel (the one-CTA Jacobi solver that replaces cuSOLVER dsyevd for n <= 32: numpy.linalg.eigh of
    huzinaga_scf.py:145,168 on the small configurations): residual, orthonormality, ascending order and the host QL
    solver's eigenvalues on random / degenerate / zero / wide-range matrices of every size 1 .. 32; only the lower
    triangle is read, like numpy.linalg.eigh."""
    out = _run_cuda_tool(tmp_path, "small_eigh_test")
    assert out.returncode == 0 and "order errors 0" in out.stdout, out.stdout[-600:] + out.stderr[-300:]


def test_block_product_kernels_against_a_plain_reference(tmp_path):
    """The subspace eigensolver's block product out = alpha (F' Y - c Y) - beta Z: the ring kernel and the single-shot
    kernel, each with and without programmatic dependent launch, against a plain reference kernel (n = 334: ragged
    last row block and contraction tail)."""
    import re

    out = _run_cuda_tool(tmp_path, "sub_apply_bench", ("334", "2", "20"))
    devs = [float(x) for x in re.findall(r"max\|out - ref\| = ([0-9.e+-]+)", out.stdout)]
    assert out.returncode == 0 and len(devs) == 8 and max(devs) < 1e-13, out.stdout[-800:] + out.stderr[-300:]


def test_huzinaga_rhf_subspace_path_with_the_single_shot_block_product(ctx):
    """Restricted (rank-2) Huzinaga SCF at n = 544: one spin per block-product launch, i.e. the shape in which the
    single-shot 8-CTA-cluster kernel is chosen on ONE GPU (on >= 2 GPUs the spin split produces it for UHF as well).
    Iterates against the oracle's full diagonalisation in every cycle."""
    n, naux, nocc, n_env = 544, 16, 7, 9
    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=4, scale=3.0 / np.sqrt(n * naux))
    b = p.cderi()
    mf = ps.DFRHF(p.ovlp, p.hcore, b, p.nelec, max_cycle=30, conv_tol=1e-8)
    tr = []
    v, g = p.v_emb[0], 2.0 * p.dm_enviro[0]
    c0, e0, d0, h0, conv0 = nr.huzinaga_scf(mf, v, g, trace=tr)
    ctx.load_cderi(b)
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, v, g, NBD_HUZINAGA)
    a0, f0, r0 = (ctx.timer_ms(k) for k in ("count:sub_applies", "count:sub_fallbacks", "count:sub_rejects"))
    c1, e1, d1, h1, info = ctx.huzinaga_scf(30, 1e-8, 1e-6, True)
    assert ctx.timer_ms("count:sub_applies") > a0 and ctx.timer_ms("count:sub_fallbacks") == f0
    assert ctx.timer_ms("count:sub_rejects") == r0
    _same_stop(info, conv0, tr)
    for k, t in enumerate(tr[: info["cycles"]]):
        assert abs(info["trace"][k, 0] - float(t["energy"])) < E_TOL, k
    assert np.abs(d1 - d0).max() < 1e-8 and np.abs(h1 - h0).max() < 1e-7 and np.abs(e1 - e0).max() < 1e-8


def test_packed_lower_fock_assembly_matches_the_square_route(ctx):
    """The multi-GPU SCF loop all-reduces the packed lower triangles of [J | K_a | K_b] and assembles the Fock matrix
    from the packed sums (scf_host.cuh: scf_build_fock).  Option packed_allreduce = 2 takes that route on one rank (the
    all-reduce is a no-op there): the iterates must be bit-identical to the square route's."""
    n, naux, nocc, n_env = 300, 20, 5, 8
    p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=6, scale=3.0 / np.sqrt(n * naux))
    ctx.load_cderi(p.cderi())
    runs = {}
    try:
        for mode in (1, 2):
            ctx.set_option("packed_allreduce", mode)
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
            runs[mode] = ctx.huzinaga_scf(12, 1e-9, 1e-7, True)
    finally:
        ctx.set_option("packed_allreduce", 1)
    a, b = runs[1], runs[2]
    assert a[4]["cycles"] == b[4]["cycles"] and np.array_equal(a[4]["trace"], b[4]["trace"])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
