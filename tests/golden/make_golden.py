"""Generates the committed fixtures of tests/golden/ (run in the DEV CONTAINER only: needs /root/reference).

    python tests/golden/make_golden.py

1. ``water_sto3g.npz``     S, hcore, exact ERI and E_nuc of the reference's test molecule (tests/molecules/water.xyz,
                           STO-3G) from oracle/gaussian_integrals.py, plus the reference's own golden energies
                           (tests/test_driver.py:56-57,76) the oracle is pinned on.
2. ``reference_runs.npz``  outputs of the UNMODIFIED reference functions, imported from /root/reference behind the
                           stub pyscf/openfermion modules of oracle/stubs.py, on seeded synthetic inputs:
                           nbed.scf.huzinaga_scf.huzinaga_scf (UHF and RHF objects, DIIS on/off, with and without
                           an initial guess), get_huzinaga_operator, energy_elec, and
                           HamiltonianBuilder._spinorb_from_spatial.
Inputs are regenerated from their seeds by the tests (nbed_b200/synthetic.py); only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nbed_b200 import synthetic as syn  # noqa: E402
from oracle import gaussian_integrals as gi  # noqa: E402
from oracle import stubs  # noqa: E402

# (name, config key, coupling scale): shared with tests/test_golden.py
SCF_CASES = [("c1", "C1_h2o_sto3g", 2.0), ("c2", "C2_h2o_ccpvdz", 3.0)]


def scf_problem(key, scale):
    cfg = dict(syn.CONFIGS[key])
    p = syn.make_problem(seed=0, scale=scale / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)
    return p, p.cderi()


def water():
    xyz = open("/root/reference/tests/molecules/water.xyz").read()
    ints = gi.integrals(gi.parse_xyz(xyz))
    np.savez_compressed(
        os.path.join(HERE, "water_sto3g.npz"), S=ints["S"], hcore=ints["T"] + ints["V"], eri=ints["eri"],
        e_nuc=ints["e_nuc"], nelectron=ints["nelectron"],
        ref_e_nuc=9.285714221677825, ref_e_uhf=-74.96099960129165,  # tests/test_driver.py:56-57
        ref_e_fci=-75.00912605315143,  # tests/test_driver.py:76
    )


def reference_runs():
    stubs.install()
    from nbed.ham_builder import HamiltonianBuilder
    from nbed.scf.embedded_hcore_funcs import energy_elec
    from nbed.scf.huzinaga_scf import get_huzinaga_operator, huzinaga_scf

    out = {}
    for name, key, scale in SCF_CASES:
        p, b = scf_problem(key, scale)
        for diis in (True, False):
            mf = stubs.make_scf("uhf", p.ovlp, p.hcore, b, p.nelec, max_cycle=40, conv_tol=1e-8)
            c, e, d, h, conv = huzinaga_scf(mf, p.v_emb, p.dm_enviro, use_DIIS=diis)
            tag = f"{name}_uhf_diis{int(diis)}"
            out.update({f"{tag}_e": e, f"{tag}_dm": np.asarray(d), f"{tag}_huz": h, f"{tag}_conv": conv,
                        f"{tag}_ncall": mf.n_jk_builds, f"{tag}_cabs": np.abs(c)})
        mf = stubs.make_scf("rhf", p.ovlp, p.hcore, b, p.nelec, max_cycle=40, conv_tol=1e-8)
        c, e, d, h, conv = huzinaga_scf(mf, p.v_emb[0], 2.0 * p.dm_enviro[0])
        tag = f"{name}_rhf"
        out.update({f"{tag}_e": e, f"{tag}_dm": np.asarray(d), f"{tag}_huz": h, f"{tag}_conv": conv,
                    f"{tag}_ncall": mf.n_jk_builds})
        # with an initial guess: the converged UHF density, perturbed
        rng = np.random.default_rng(5)
        r = rng.normal(size=(2, p.n, p.n)) * 1e-3
        dm0 = out[f"{name}_uhf_diis1_dm"] + r + r.transpose(0, 2, 1)
        mf = stubs.make_scf("uhf", p.ovlp, p.hcore, b, p.nelec, max_cycle=40, conv_tol=1e-8)
        c, e, d, h, conv = huzinaga_scf(mf, p.v_emb, p.dm_enviro, dm_initial_guess=dm0)
        tag = f"{name}_uhf_guess"
        out.update({f"{tag}_e": e, f"{tag}_dm": np.asarray(d), f"{tag}_huz": h, f"{tag}_conv": conv,
                    f"{tag}_ncall": mf.n_jk_builds})
        # energy_elec with a spin-resolved core Hamiltonian (embedded_hcore_funcs.py:11-46)
        mf = stubs.make_scf("uhf", p.ovlp, p.hcore, b, p.nelec)
        h3 = p.hcore + p.v_emb
        dmc = out[f"{name}_uhf_diis1_dm"]
        out[f"{name}_energy_elec"] = np.array(energy_elec(mf, dmc, h3, None))
    # get_huzinaga_operator on random operands, rank 3 and rank 2 (huzinaga_scf.py:65-90)
    rng = np.random.default_rng(11)
    f3, g3 = rng.normal(size=(2, 9, 9)), rng.normal(size=(2, 9, 9))
    out["huzop_rank3"] = get_huzinaga_operator(f3, g3, np.zeros_like(g3))
    out["huzop_rank2"] = get_huzinaga_operator(f3[0], g3[0], np.zeros_like(g3[0]))
    # _spinorb_from_spatial (ham_builder.py:158-216) incl. values around the EQ_TOLERANCE cliff
    one = rng.normal(size=(2, 3, 3))
    two = rng.normal(size=(4, 3, 3, 3, 3))
    two[0, 0, 1, 2, 0] = 0.99e-8
    two[2, 2, 1, 0, 1] = -1.01e-8
    one[1, 2, 0] = 5e-9
    h1, h2 = HamiltonianBuilder._spinorb_from_spatial(None, one, two)
    out.update({"spinorb_h1": h1, "spinorb_h2": h2})
    np.savez_compressed(os.path.join(HERE, "reference_runs.npz"), **out)
    return out


def ks_problem():
    """water / STO-3G with B3LYP on a fixed Becke grid: global UKS, then the O 1s orbital as frozen environment with
    the DFT-in-DFT embedding potential veff[g_act + g_env] - veff[g_act] (nbed/driver.py:847-849).  Shared with the tests."""
    import scipy.linalg

    from oracle import gto_restatement as g
    from oracle import pyscf_restatement as ps
    from oracle import xc_restatement as xc

    atoms = g.parse_xyz(open(os.path.join(HERE, "water.xyz")).read())
    w = np.load(os.path.join(HERE, "water_sto3g.npz"))
    s, h = w["S"], w["hcore"]
    cd = ps.cholesky_eri_exact(w["eri"])
    atm, bas, env = g.make_env(atoms, gi.STO3G)
    coords, wts = xc.becke_grid(atoms, 50, 14, 28)
    ao = xc.eval_ao(g.shells_from_env(atm, bas, env), coords)
    mf = xc.DFUKS(s, h, cd, (5, 5), ao, wts, "b3lyp", e_nuc=float(w["e_nuc"]), max_cycle=50, conv_tol=1e-10)
    _, c = scipy.linalg.eigh(h, s)
    dm0 = np.array([c[:, :5] @ c[:, :5].T] * 2)
    conv, e_tot, _, c0, _ = ps.scf_kernel(mf, conv_tol=1e-10, dm0=dm0)
    c_env = np.array([c0[0][:, :1], c0[1][:, :1]])
    dm_env = c_env @ c_env.swapaxes(-1, -2)
    dm_act = np.array([c0[k][:, 1:5] @ c0[k][:, 1:5].T for k in range(2)])
    v_emb = np.asarray(mf.get_veff(dm=dm_act + dm_env)) - np.asarray(mf.get_veff(dm=dm_act))
    return dict(atoms=atoms, s=s, h=h, cderi=cd, basis=(atm, bas, env), coords=coords, weights=wts, ao=ao, dm0=dm0,
                e_nuc=float(w["e_nuc"]), global_ks=mf, global_conv=conv, global_e_tot=e_tot, c_env=c_env, dm_env=dm_env,
                v_emb=v_emb)


def reference_runs_ks():
    """The Kohn-Sham branch of the UNMODIFIED reference loop (huzinaga_scf.py:176-180, calculate_ks_energy :36-62) over
    the stub UKS object of oracle/stubs.py (DF J/K + the XC restatement) -> reference_runs_ks.npz."""
    stubs.install()
    from nbed.scf.huzinaga_scf import huzinaga_scf

    p = ks_problem()
    act = stubs.make_scf("uks", p["s"], p["h"], p["cderi"], (4, 4), ao=p["ao"], weights=p["weights"], xc="b3lyp",
                         max_cycle=40, conv_tol=1e-9)
    c, e, d, hz, conv = huzinaga_scf(act, p["v_emb"], p["dm_env"], dm_conv_tol=1e-7)
    # the same embedding driven through a restricted Kohn-Sham object (rank-2 arrays, doubled environment density:
    # occupied/base.py:84-85), the object type of the reference's tests/test_scf.py:19-40
    rks = stubs.make_scf("rks", p["s"], p["h"], p["cderi"], (4, 4), ao=p["ao"], weights=p["weights"], xc="b3lyp",
                         max_cycle=40, conv_tol=1e-9)
    from nbed.scf.huzinaga_scf import calculate_ks_energy  # (nbed.scf.huzinaga_scf the attribute is the function)
    cr, er, dr, hzr, convr = huzinaga_scf(rks, p["v_emb"][0], 2.0 * p["dm_env"][0], dm_conv_tol=1e-7)
    out_rks = dict(rks_e=er, rks_dm=np.asarray(dr), rks_huz=hzr, rks_conv=convr, rks_n_veff=rks.n_xc_builds,
                   rks_energy=float(calculate_ks_energy(rks, p["v_emb"][0], dr, hzr)))
    out = dict(global_e_tot=p["global_e_tot"], global_energy_elec=np.array(p["global_ks"].energy_elec()),
               ref_global_e_tot=-75.3091447400438, ref_global_energy_elec=np.array([-84.59485896172163, 37.93302591280513]),
               v_emb=p["v_emb"], c_env=p["c_env"], ks_e=e, ks_dm=np.asarray(d), ks_huz=hz, ks_conv=conv,
               ks_n_veff=act.n_xc_builds, ngrid=len(p["weights"]), **out_rks)
    np.savez_compressed(os.path.join(HERE, "reference_runs_ks.npz"), **out)
    return out


if __name__ == "__main__":
    water()
    ok = reference_runs_ks()
    print("KS fixture: global E =", ok["global_e_tot"], "reference golden", ok["ref_global_e_tot"], "veff builds", ok["ks_n_veff"])
    o = reference_runs()
    print("written:", sorted(os.listdir(HERE)), len(o), "arrays")
