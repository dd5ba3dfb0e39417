"""NumPy restatement of the reference's own hot-path functions.  TEST INFRASTRUCTURE.

Every function cites the reference file:line it follows (paths relative to /root/reference).
The SCF object passed in is any object satisfying the protocol of SURVEY.md §8(b); the oracle uses
``oracle.pyscf_restatement.DFUHF`` / ``DFRHF``.  Validated against the reference's unmodified code by ``tests/golden/make_golden.py`` /
``make_golden_fullsize.py`` (dev container: they import /root/reference behind oracle/stubs.py) and frozen in
``tests/golden``; ``tests/test_golden.py`` holds the restatement to those outputs.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from . import pyscf_restatement as ps

EQ_TOLERANCE = 1e-8  # openfermion.config.EQ_TOLERANCE (openfermion 1.7.1), used at nbed/ham_builder.py:213-214


# ---- nbed/scf/huzinaga_scf.py:65-90 ---------------------------------------------------------
def get_huzinaga_operator(fock, dm_occ_S, dm_virt_S):
    fds_occ = np.einsum("...ij,...jk->...ik", fock, dm_occ_S)
    huz_occ = fds_occ + np.swapaxes(fds_occ, -1, -2)
    huz_occ *= (-0.5) if fds_occ.ndim == 2 else (-1.0)
    fds_virt = np.einsum("...ij,...jk->...ik", fock, dm_virt_S)
    huz_virt = (
        fds_virt
        + np.swapaxes(fds_virt, -1, -2)
        - 2 * np.einsum("...ij,...jk->...ik", np.swapaxes(dm_virt_S, -1, -2), fds_virt)
    )
    huz_virt *= (-0.5) if fds_virt.ndim == 2 else (-1.0)
    return huz_occ + huz_virt


# ---- nbed/scf/huzinaga_scf.py:93-206 --------------------------------------------------------
def huzinaga_scf(
    scf_method,
    embedding_potential,
    dm_environment_occupied,
    dm_environment_virtual=None,
    dm_conv_tol=1e-6,
    dm_initial_guess=None,
    use_DIIS=True,
    trace=None,
):
    """Same iterate sequence as the reference loop, Hartree-Fock objects and - ``scf_method.is_ks`` - Kohn-Sham objects
    (the isinstance test of :176 selects calculate_ks_energy, :36-62, which calls get_veff a second time per cycle).

    ``trace`` (list) receives per-cycle dicts {energy, norm_dm_diff} for iterate-level parity tests.
    """
    s_mat = scf_method.get_ovlp()  # :126
    s_neg_half = scipy.linalg.fractional_matrix_power(s_mat, -0.5)  # :128
    adiis = ps.DIIS() if use_DIIS else None  # :130
    dm_occ_S = np.einsum("...ij,jk->...ik", dm_environment_occupied, s_mat)  # :132
    if dm_environment_virtual is not None:
        dm_virt_S = np.einsum("...ij,jk->...ik", dm_environment_virtual, s_mat)
    else:
        dm_virt_S = np.zeros(dm_occ_S.shape)  # :136
    if dm_initial_guess is None:  # :139-148
        fock = scf_method.get_hcore() + embedding_potential
        fock = fock + get_huzinaga_operator(fock, dm_occ_S, dm_virt_S)
        fock_ortho = s_neg_half @ fock @ s_neg_half
        mo_energy, mo_coeff_ortho = np.linalg.eigh(fock_ortho)
        mo_coeff_std = s_neg_half @ mo_coeff_ortho
        mo_occ = scf_method.get_occ(mo_energy, mo_coeff_std)
        dm_initial_guess = scf_method.make_rdm1(mo_coeff=mo_coeff_std, mo_occ=mo_occ)
    density_matrix = dm_initial_guess
    conv_flag = False
    scf_energy_prev = 0
    for i in range(scf_method.max_cycle):  # :154
        vhf = scf_method.get_veff(dm=density_matrix)  # :156
        fock = scf_method.get_hcore() + embedding_potential + vhf  # :157
        huzinaga_op = get_huzinaga_operator(fock, dm_occ_S, dm_virt_S)  # :159
        fock = fock + huzinaga_op  # :160
        if use_DIIS and (i > 1):  # :162
            fock = adiis.update(fock)
        fock_ortho = s_neg_half @ fock @ s_neg_half  # :166
        mo_energy, mo_coeff_ortho = np.linalg.eigh(fock_ortho)  # :168
        mo_coeff_std = s_neg_half @ mo_coeff_ortho
        mo_occ = scf_method.get_occ(mo_energy, mo_coeff_std)
        dm_mat_old = density_matrix
        density_matrix = scf_method.make_rdm1(mo_coeff=mo_coeff_std, mo_occ=mo_occ)  # :174
        if getattr(scf_method, "is_ks", False):  # :176-180 -> calculate_ks_energy, :55-61
            vhf_updated = scf_method.get_veff(dm=density_matrix)
            scf_energy = vhf_updated.ecoul + vhf_updated.exc
            scf_energy = scf_energy + np.einsum(
                "...ij, ...ji->...", density_matrix, (scf_method.get_hcore() + huzinaga_op + embedding_potential))
        else:
            hamiltonian = scf_method.get_hcore() + embedding_potential + 0.5 * vhf + huzinaga_op  # :182-184
            scf_energy = np.einsum("...ij,...ji->...", hamiltonian, density_matrix)  # :185
        run_diff = np.max(np.abs(scf_energy - scf_energy_prev))  # :191
        norm_dm_diff = np.max(np.linalg.norm(density_matrix - dm_mat_old, axis=(-2, -1)))  # :192-194
        if trace is not None:
            trace.append({"energy": np.array(scf_energy, dtype=float), "norm_dm_diff": float(norm_dm_diff)})
        if (run_diff < scf_method.conv_tol) and (norm_dm_diff < dm_conv_tol):  # :196
            conv_flag = True
            break
        scf_energy_prev = scf_energy
    return mo_coeff_std, mo_energy, density_matrix, huzinaga_op, conv_flag


# ---- nbed/scf/embedded_hcore_funcs.py:11-46 -------------------------------------------------
def energy_elec(mf, dm=None, h1e=None, vhf=None):
    if dm is None:
        dm = mf.make_rdm1()
    if h1e is None:
        h1e = mf.get_hcore()
    if isinstance(dm, np.ndarray) and dm.ndim == 2:
        dm = np.array((dm * 0.5, dm * 0.5))
    if vhf is None:
        vhf = mf.get_veff(mf.mol, dm)
    e1 = np.einsum("ij,ji->", h1e[0], dm[0])
    e1 += np.einsum("ij,ji->", h1e[1], dm[1])
    e_coul = (np.einsum("ij,ji->", vhf[0], dm[0]) + np.einsum("ij,ji->", vhf[1], dm[1])) * 0.5
    e_elec = (e1 + e_coul).real
    mf.scf_summary["e1"] = e1.real
    mf.scf_summary["e2"] = e_coul.real
    return e_elec, e_coul


# ---- nbed/driver.py:433-449 -----------------------------------------------------------------
def env_projector(s_mat, dm_enviro):
    p_alpha = s_mat @ dm_enviro[0] @ s_mat
    if dm_enviro.ndim == 2:
        return p_alpha
    return np.array([p_alpha, s_mat @ dm_enviro[1] @ s_mat])


# ---- nbed/driver.py:500-538 (+ PySCF kernel(), restated in pyscf_restatement.scf_kernel) ------
def mu_embed(localized_scf, embedding_potential, dm_enviro, mu_level_shift=1e6, dm0=None, trace=None):
    v_emb = mu_level_shift * env_projector(localized_scf.get_ovlp(), dm_enviro) + embedding_potential  # :518
    if v_emb.ndim == 3:
        localized_scf.energy_elec = lambda *args: energy_elec(localized_scf, *args)  # :521-522
    hcore_std = localized_scf.get_hcore
    localized_scf.get_hcore = lambda *args: hcore_std(*args) + v_emb  # :529
    ps.scf_kernel(localized_scf, conv_tol=localized_scf.conv_tol, dm0=dm0, trace=trace)  # :533
    return localized_scf, v_emb


# ---- nbed/driver.py:540-632 -----------------------------------------------------------------
def huzinaga_embed(active_scf, embedding_potential, dm_enviro, dmat_initial_guess=None, trace=None):
    c, e, dm, huz, conv = huzinaga_scf(
        active_scf, embedding_potential, dm_enviro, dm_environment_virtual=None, dm_conv_tol=1e-6,
        dm_initial_guess=dmat_initial_guess, trace=trace,
    )  # :576-589
    hcore_std = active_scf.get_hcore()
    v_emb = huz + embedding_potential  # :596
    active_scf.get_hcore = lambda *args: hcore_std + v_emb  # :597
    if np.ndim(dm_enviro) == 3:
        active_scf.energy_elec = lambda *args: energy_elec(active_scf, *args)  # :599-600
    active_scf.mo_occ = active_scf.get_occ(e, c)  # :602
    active_scf.mo_coeff = c  # :621
    active_scf.mo_energy = e
    active_scf.e_tot = active_scf.energy_tot(dm=dm)  # :627
    active_scf.converged = conv
    return active_scf, v_emb


# ---- nbed/ham_builder.py:53-96 --------------------------------------------------------------
def one_body_integrals(scf_method, restricted):
    c = scf_method.mo_coeff
    hcore = scf_method.get_hcore()
    if hcore.ndim == 2:
        hcore = [hcore] * 2
    if not restricted:
        return np.array([c[0].T @ hcore[0] @ c[0], c[1].T @ hcore[1] @ c[1]])
    return np.array([c.T @ scf_method.get_hcore() @ c] * 2)


# ---- nbed/ham_builder.py:98-156 -------------------------------------------------------------
def two_body_integrals(cderi, mo_coeff, restricted):
    """(4, m, m, m, m), [blk][p,r,s,q] = (pq|rs); blocks aaaa, bbbb, aabb, bbaa (:119-124)."""
    if not restricted:
        ca, cb = mo_coeff[0], mo_coeff[1]
        if ca.shape[1] != cb.shape[1]:
            raise ValueError("Must localize the same number of alpha and beta orbitals.")  # :109-112
        m = ca.shape[1]
        out = []
        for cs in ((ca, ca, ca, ca), (cb, cb, cb, cb), (ca, ca, cb, cb), (cb, cb, ca, ca)):
            eri = ps.ao2mo_restore(1, ps.ao2mo_kernel(cderi, cs), m)  # :128-131
            out.append(np.asarray(eri.transpose(0, 2, 3, 1), order="C"))  # :132
        return np.stack(out, axis=0)
    m = mo_coeff.shape[1]
    eri = ps.ao2mo_restore(1, ps.ao2mo_kernel(cderi, mo_coeff), m)
    return np.stack([np.asarray(eri.transpose(0, 2, 3, 1), order="C")] * 4, axis=0)  # :152-154


# ---- nbed/ham_builder.py:158-216 (vectorised; the reference's 4-deep Python loop) ------------
def spinorb_from_spatial(one_body_integrals_, two_body_integrals_):
    m = one_body_integrals_[0].shape[0]
    nq = 2 * m
    h1 = np.zeros((nq, nq))
    h2 = np.zeros((nq, nq, nq, nq))
    h1[0::2, 0::2] = one_body_integrals_[0]
    h1[1::2, 1::2] = one_body_integrals_[1]
    h2[0::2, 0::2, 0::2, 0::2] = two_body_integrals_[0]  # aaaa  :190-192
    h2[1::2, 1::2, 1::2, 1::2] = two_body_integrals_[1]  # bbbb  :194-196
    h2[0::2, 1::2, 1::2, 0::2] = two_body_integrals_[2]  # abba  :200-202
    h2[1::2, 0::2, 0::2, 1::2] = two_body_integrals_[3]  # baab  :204-206
    h1[np.absolute(h1) < EQ_TOLERANCE] = 0.0  # :213
    h2[np.absolute(h2) < EQ_TOLERANCE] = 0.0  # :214
    return h1, h2


# ---- nbed/ham_builder.py:218-254 ------------------------------------------------------------
def build_hamiltonian(scf_method, cderi, constant_e_shift=0.0, restricted=False):
    one = one_body_integrals(scf_method, restricted)
    two = two_body_integrals(cderi, scf_method.mo_coeff, restricted)
    h1, h2 = spinorb_from_spatial(one, two)
    return constant_e_shift, h1, 0.5 * h2
