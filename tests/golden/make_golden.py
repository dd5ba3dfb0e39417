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


if __name__ == "__main__":
    water()
    o = reference_runs()
    print("written:", sorted(os.listdir(HERE)), len(o), "arrays")
