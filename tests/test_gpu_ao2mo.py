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
