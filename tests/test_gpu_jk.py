"""GPU parity: 3-centre tensor layout and the density-fitted J/K build against the CPU oracle."""
import os

import numpy as np
import pytest

from nbed_b200 import synthetic as syn
from oracle import pyscf_restatement as ps

pytestmark = pytest.mark.gpu

SIZES = [(7, 21, (4, 4)), (24, 72, (5, 4)), (33, 50, (3, 0)), (45, 64, (9, 11)), (174, 96, (9, 9)), (100, 40, (17, 20)),
         # more than 8 AO panels: several accumulator slots per consumer warp (NSLOT = 2, 4, 6) and both column widths
         (300, 24, (5, 5)), (300, 9, (3, 4)), (520, 12, (3, 0)), (1000, 6, (5, 4)), (1376, 6, (5, 5)), (1376, 4, (3, 3)),
         # 9 / 10 trailing columns: 8 DMMA columns + 1-2 columns on the FMA pipe, alone and behind 16-column slices
         (100, 40, (13, 12)), (300, 7, (5, 4)), (520, 8, (13, 13)), (1376, 5, (5, 4)), (1376, 3, (13, 13))]


def _cderi(n, naux, seed=3):
    return syn.synth_cderi_rows(seed, n, 0.7 / np.sqrt(n * naux) * 3, np.arange(naux))


@pytest.mark.parametrize("n,naux", [(7, 21), (24, 72), (33, 5), (64, 3), (65, 9), (174, 40)])
def test_cderi_roundtrip_bit_exact(ctx, n, naux):
    rng = np.random.default_rng(n)
    b = rng.normal(size=(naux, n * (n + 1) // 2))
    ctx.load_cderi(b)
    back = ctx.cderi_download(0, naux)
    assert np.array_equal(back, b)
    part = ctx.cderi_download(naux // 2, naux - naux // 2)
    assert np.array_equal(part, b[naux // 2 :])


@pytest.mark.parametrize("n,naux", [(7, 21), (45, 17), (174, 12)])
def test_cderi_synth_bit_exact(ctx, n, naux):
    ctx.cderi_alloc(n, naux)
    ctx.cderi_synth(5, 0.125, 11)
    want = syn.synth_cderi_rows(5, n, 0.125, np.arange(11, 11 + naux))
    assert np.array_equal(ctx.cderi_download(0, naux), want)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,naux,nocc", SIZES)
def test_jk_occupied_orbitals(ctx, n, naux, nocc, variant):
    rng = np.random.default_rng(n + naux)
    b = _cderi(n, naux)
    ctx.load_cderi(b)
    ctx.set_option("jk_variant", variant)
    ctx.set_option("gemm_variant", variant)
    try:
        orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
        vj, vk = ctx.jk_orbitals(orbs)
    finally:
        ctx.set_option("jk_variant", 0)
        ctx.set_option("gemm_variant", 0)
    rj, rk = ps.df_get_jk_occ(b, orbs)
    scale = max(1.0, np.abs(rj).max(), np.abs(rk).max())
    assert np.abs(vj - rj).max() <= 1e-12 * scale
    assert np.abs(vk - rk).max() <= 1e-12 * scale
    assert np.array_equal(vj, vj.transpose(0, 2, 1))
    assert np.array_equal(vk, vk.transpose(0, 2, 1))


def test_jk_aux_chunking(ctx):
    n, naux, nocc = 45, 64, (9, 11)
    rng = np.random.default_rng(1)
    b = _cderi(n, naux)
    ctx.load_cderi(b)
    orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
    vj0, vk0 = ctx.jk_orbitals(orbs)
    ctx.set_option("x_budget_mb", 0)  # one aux row per chunk
    try:
        vj1, vk1 = ctx.jk_orbitals(orbs)
    finally:
        ctx.set_option("x_budget_mb", 3072)
    assert np.abs(vj0 - vj1).max() < 1e-13
    assert np.abs(vk0 - vk1).max() < 1e-13


@pytest.mark.parametrize("n,naux", [(7, 21), (45, 30), (100, 16)])
def test_jk_dense_density(ctx, n, naux):
    rng = np.random.default_rng(n)
    b = _cderi(n, naux)
    ctx.load_cderi(b)
    a = rng.normal(size=(2, n, n)) / n
    dm = a + a.transpose(0, 2, 1)  # indefinite symmetric "densities"
    vj, vk = ctx.jk_dm(dm)
    rj, rk = ps.df_get_jk(b, dm)
    scale = max(1.0, np.abs(rj).max(), np.abs(rk).max())
    assert np.abs(vj - rj).max() <= 1e-11 * scale
    assert np.abs(vk - rk).max() <= 1e-11 * scale


def test_jk_empty_inputs(ctx):
    n = 12
    ctx.cderi_alloc(n, 0)  # rank with an empty aux shard
    vj, vk = ctx.jk_orbitals([np.ones((n, 2)) / n, np.ones((n, 1))])
    assert not vj.any() and not vk.any()
    b = _cderi(n, 8)
    ctx.load_cderi(b)
    vj, vk = ctx.jk_orbitals([np.zeros((n, 0)), np.ones((n, 1)) / n])
    rj, rk = ps.df_get_jk_occ(b, [np.zeros((n, 0)), np.ones((n, 1)) / n])
    assert np.abs(vj - rj).max() < 1e-13 and np.abs(vk - rk).max() < 1e-13
    assert not vk[0].any() and not vj[0].any()


def test_jk_and_ao2mo_fuzz_small_shapes(ctx):
    """Seeded sweep over awkward shapes: tile-boundary AO counts, tiny aux ranges, empty / full / unequal orbital sets."""
    from oracle import nbed_restatement as nr

    rng = np.random.default_rng(2024)
    shapes = [(1, 1), (2, 3), (31, 2), (32, 5), (33, 4), (63, 3), (64, 2), (65, 7), (96, 1), (97, 3), (129, 2), (255, 2),
              (256, 3), (257, 2)]
    for n, naux in shapes:
        b = rng.normal(size=(naux, n * (n + 1) // 2)) / n
        ctx.load_cderi(b)
        for nocc in {(min(n, 1), 0), (min(n, 3), min(n, 2)), (min(n, 9), min(n, 16))}:
            orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
            vj, vk = ctx.jk_orbitals(orbs)
            rj, rk = ps.df_get_jk_occ(b, orbs)
            scale = max(1.0, np.abs(rj).max(), np.abs(rk).max())
            assert np.abs(vj - rj).max() <= 1e-12 * scale, (n, naux, nocc)
            assert np.abs(vk - rk).max() <= 1e-12 * scale, (n, naux, nocc)
        m = min(n, 5)
        c = rng.normal(size=(2, n, m)) / np.sqrt(n)
        got = ctx.ao2mo(c[0], c[1])
        want = nr.two_body_integrals(b, c, restricted=False)
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (n, naux)


def test_jk_overlap_modes_are_bit_identical(ctx):
    """Pass 2 behind the Gram (0), next to it on the side stream (1, default) or launched in front of it with
    programmatic stream serialization (3): scheduling only, the sums are taken in the same order."""
    n, naux, nocc = 300, 24, (5, 5)
    rng = np.random.default_rng(11)
    ctx.load_cderi(_cderi(n, naux))
    orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
    out = {}
    try:
        for mode in (1, 0, 3):
            ctx.set_option("overlap", mode)
            out[mode] = ctx.jk_orbitals(orbs)
    finally:
        ctx.set_option("overlap", 1)
    for mode in (0, 3):
        assert np.array_equal(out[mode][0], out[1][0]) and np.array_equal(out[mode][1], out[1][1]), mode


def test_jk_hybrid_panel_matches_padded_panel(ctx):
    """9 / 10 column slices: 8 DMMA columns + 1-2 FMA-pipe columns (default) against the same slices padded to 16 DMMA
    columns (panel_hybrid = 0).  The FMA columns sum in a different order, so agreement is to rounding, not bitwise."""
    for n, naux, nocc in [(300, 12, (5, 5)), (1376, 3, (5, 4)), (520, 6, (13, 12))]:
        rng = np.random.default_rng(n)
        ctx.load_cderi(_cderi(n, naux))
        orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
        vj1, vk1 = ctx.jk_orbitals(orbs)
        ctx.set_option("panel_hybrid", 0)
        try:
            vj0, vk0 = ctx.jk_orbitals(orbs)
        finally:
            ctx.set_option("panel_hybrid", 1)
        scale = max(1.0, np.abs(vk0).max(), np.abs(vj0).max())
        assert np.abs(vj1 - vj0).max() <= 1e-13 * scale and np.abs(vk1 - vk0).max() <= 1e-13 * scale


@pytest.mark.parametrize("n,naux,nocc", [(520, 12, (5, 5)), (1376, 7, (5, 4)), (1000, 9, (3, 0)), (640, 40, (13, 13))])
def test_k_gram_stream_k_kernel(ctx, n, naux, nocc):
    """The exchange Gram as the stream-K symmetric rank-k kernel (syrk.cuh; n >= 512, even) against the oracle and
    against the generic lower-tile GEMM, for grids that cut every tile (one CTA per SM), that hold whole tiles per CTA
    (7 CTAs: the direct-write path) and that have more CTAs than k-steps per tile."""
    rng = np.random.default_rng(n + naux)
    b = _cderi(n, naux)
    ctx.load_cderi(b)
    orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
    rj, rk = ps.df_get_jk_occ(b, orbs)
    scale = max(1.0, np.abs(rj).max(), np.abs(rk).max())
    try:
        ctx.set_option("syrk", 0)
        _, vk_gemm = ctx.jk_orbitals(orbs)
        ctx.set_option("syrk", 2)  # also next to pass 2 on shared SMs (1 = only alone or on its own SMs)
        ctx.set_option("pair_split", 0)
        for ctas in (0, 7, 1000):
            ctx.set_option("syrk_ctas", ctas)
            vj, vk = ctx.jk_orbitals(orbs)
            assert np.abs(vk - rk).max() <= 1e-12 * scale, ctas
            assert np.abs(vj - rj).max() <= 1e-12 * scale, ctas
            assert np.abs(vk - vk_gemm).max() <= 1e-13 * scale, ctas
            assert np.array_equal(vk, vk.transpose(0, 2, 1))
    finally:
        ctx.set_option("syrk", 1)
        ctx.set_option("syrk_ctas", 0)
        ctx.set_option("pair_split", -1)


def test_pair_modes_agree_and_repeat(ctx):
    """Pass 2 next to the stream-K Gram: shared SMs, spatial split (automatic and forced), serial.  Every mode must agree
    with the serial result to rounding and repeat BIT-identically (regression test for the ring stage that was released
    before its loads had returned, profiles/r02_pass2_race.md)."""
    n, naux, nocc = 1376, 192, (5, 5)
    rng = np.random.default_rng(5)
    ctx.cderi_alloc(n, naux)
    ctx.cderi_synth(3, 3.0 / np.sqrt(n * naux), 0)
    orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
    try:
        ctx.set_option("overlap", 0)
        j0, k0 = ctx.jk_orbitals(orbs)
        ctx.set_option("overlap", 1)
        scale = np.abs(k0).max()
        for split, syrk in ((-1, 1), (0, 2), (0, 1), (60, 1), (100, 1)):
            ctx.set_option("pair_split", split)
            ctx.set_option("syrk", syrk)
            j1 = k1 = None
            for rep in range(5):
                j, k = ctx.jk_orbitals(orbs)
                if rep == 0:
                    j1, k1 = j, k
                # (the number of P-ranges of pass 2 follows its CTA count, so modes differ from each other by rounding)
                assert np.array_equal(j, j1) and np.array_equal(k, k1), (split, syrk, rep, np.abs(j - j1).max())
                assert np.abs(j - j0).max() <= 1e-14 * np.abs(j0).max(), (split, syrk, rep)
                assert np.abs(k - k0).max() <= 1e-13 * scale, (split, syrk, rep)
    finally:
        ctx.set_option("overlap", 1)
        ctx.set_option("pair_split", -1)
        ctx.set_option("syrk", 1)
