"""GPU parity: active-space AO->MO transform, one-body transform and the spin-orbital scatter."""
import numpy as np
import pytest

from nbed_b200 import synthetic as syn
from oracle import nbed_restatement as nr

pytestmark = pytest.mark.gpu

TOL = 1e-10  # per element, BASELINE.json north_star


def _setup(n, naux, m, seed=0):
    p = syn.make_problem(n=n, naux=naux, nocc=2, n_env=1, seed=seed, scale=2.0 / np.sqrt(n * naux))
    return p, p.cderi(), syn.random_orthonormal_mos(p.ovlp, m, seed)


@pytest.mark.parametrize("n,naux,m", [(7, 21, 5), (24, 72, 12), (45, 33, 9), (174, 60, 24), (70, 40, 40)])
def test_ao2mo_unrestricted(ctx, n, naux, m):
    p, b, c = _setup(n, naux, m)
    ctx.load_cderi(b)
    got = ctx.ao2mo(c[0], c[1])
    want = nr.two_body_integrals(b, c, restricted=False)
    assert got.shape == (4, m, m, m, m)
    assert np.abs(got - want).max() < TOL * max(1.0, np.abs(want).max())
    # bbaa is the transpose of aabb in chemist order: out[3][p,r,s,q] = out[2][r,p,q,s]
    assert np.abs(got[3] - got[2].transpose(1, 0, 3, 2)).max() < 1e-13


def test_ao2mo_restricted(ctx):
    p, b, c = _setup(24, 72, 10)
    ctx.load_cderi(b)
    got = ctx.ao2mo(c[0])
    want = nr.two_body_integrals(b, c[0], restricted=True)
    assert np.abs(got - want).max() < TOL * max(1.0, np.abs(want).max())


def test_one_body_and_spinorb(ctx):
    p, b, c = _setup(24, 72, 6)
    ctx.load_cderi(b)
    h3 = np.array([p.hcore + p.v_emb[0], p.hcore + p.v_emb[1]])
    one = ctx.one_body(h3, c[0], c[1])
    want1 = np.array([c[s].T @ h3[s] @ c[s] for s in range(2)])
    assert np.abs(one - want1).max() < 1e-12
    one2 = ctx.one_body(p.hcore, c[0])
    assert np.abs(one2[0] - c[0].T @ p.hcore @ c[0]).max() < 1e-12 and np.array_equal(one2[0], one2[1])
    two = ctx.ao2mo(c[0], c[1])
    two[0, 0, 1, 2, 3] = 0.9e-8  # straddle the EQ_TOLERANCE cliff
    two[1, 3, 2, 1, 0] = -1.1e-8
    h1, h2 = ctx.spinorb_from_spatial(one, two, eq_tol=1e-8, two_body_scale=0.5)
    r1, r2 = nr.spinorb_from_spatial(one, two)
    assert np.array_equal(h1, r1)
    assert np.array_equal(h2, 0.5 * r2)
    assert h2[0, 2, 4, 6] == 0.0 and h2[7, 5, 3, 1] == -0.55e-8


def test_build_hamiltonian_fused_and_pauli_term_set(ctx):
    """HamiltonianBuilder.build() (one fused device call) against the oracle, down to the qubit Hamiltonian: the
    Jordan-Wigner Pauli-term SET must be identical and the coefficients equal to 1e-10 (BASELINE.json north star;
    openfermion is absent, oracle/fock_space.py enumerates the terms)."""
    from nbed_b200 import B200RHF, B200UHF, HamiltonianBuilder
    from oracle import fock_space as fs
    from oracle import pyscf_restatement as ps

    p, b, c = _setup(24, 72, 3, seed=4)
    ctx.load_cderi(b)
    h3 = np.array([p.hcore + p.v_emb[0], p.hcore + p.v_emb[1]])
    mf = B200UHF(ctx, p.ovlp, p.hcore, (2, 2))
    mf.get_hcore = lambda *a: h3  # the embedded (spin-resolved) core Hamiltonian, as patched by the driver
    mf.mo_coeff, mf.mo_occ = c, np.array([[1, 1, 0], [1, 1, 0]], dtype=float)
    const, h1, h2 = HamiltonianBuilder(mf, constant_e_shift=-1.25).build()
    ref = ps.DFUHF(p.ovlp, p.hcore, b, (2, 2))
    ref.get_hcore = lambda *a: h3
    ref.mo_coeff = c
    rconst, r1, r2 = nr.build_hamiltonian(ref, b, -1.25, restricted=False)
    assert const == rconst and h1.shape == (6, 6) and h2.shape == (6,) * 4
    assert np.abs(h1 - r1).max() < TOL and np.abs(h2 - r2).max() < TOL
    assert np.array_equal(h1 == 0.0, r1 == 0.0) and np.array_equal(h2 == 0.0, r2 == 0.0)  # same truncation pattern
    got, want = fs.pauli_terms(const, h1, h2), fs.pauli_terms(rconst, r1, r2)
    assert set(got) == set(want) and len(got) > 50
    assert max(abs(got[k] - want[k]) for k in want) < TOL
    # unfused path gives the same tensors
    one, two = ctx.one_body(h3, c[0], c[1]), ctx.ao2mo(c[0], c[1])
    u1, u2 = ctx.spinorb_from_spatial(one, two, 1e-8, 0.5)
    assert np.array_equal(u1, h1) and np.array_equal(u2, h2)
    # restricted object: the four blocks are copies of one (ham_builder.py:152-154)
    mr = B200RHF(ctx, p.ovlp, p.hcore, (2, 2))
    mr.mo_coeff, mr.mo_occ = c[0], np.array([2.0, 2.0, 0.0])
    _, q1, q2 = HamiltonianBuilder(mr).build()
    rr = ps.DFRHF(p.ovlp, p.hcore, b, (2, 2))
    rr.mo_coeff = c[0]
    _, s1, s2 = nr.build_hamiltonian(rr, b, 0.0, restricted=True)
    assert np.abs(q1 - s1).max() < TOL and np.abs(q2 - s2).max() < TOL
