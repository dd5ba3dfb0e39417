"""Exchange-correlation on an atom-centred grid: restatement of what the reference reaches through
``pyscf.dft`` (numint + libxc) on the Kohn-Sham branch of the hot path.  TEST INFRASTRUCTURE.

Reference call sites: ``scf_method.get_veff(dm=...)`` on RKS/UKS objects (nbed/scf/huzinaga_scf.py:55,156), the
``.ecoul`` / ``.exc`` tags it reads (:56), ``calculate_ks_energy`` (:36-62), ``_init_local_ks`` (nbed/driver.py:289-313),
and the global ``dft.RKS/UKS(...).kernel()`` (:163-181).  PySCF and libxc are not in this image, so the pieces are
restated from their published definitions:

* ``b3lyp``      libxc HYB_GGA_XC_B3LYP (what pyscf 2.9.0 resolves 'b3lyp' to): 0.08 Slater + 0.72 B88 + 0.19 VWN(RPA) +
                 0.81 LYP + 0.20 exact exchange.  Energy density as a function of (rho_a, rho_b, sigma_aa, sigma_ab,
                 sigma_bb); first derivatives by forward-mode automatic differentiation (``Jet``) so that no hand-derived
                 derivative formula can disagree with the energy expression.
* ``nr_uks``     pyscf.dft.numint.NumInt.nr_uks for LDA/GGA: densities and gradients from AO values, exc, vxc matrices.
* ``DFUKS``      pyscf.dft.uks.UKS.get_veff / energy_elec over the density-fitted J/K of pyscf_restatement.
* ``becke_grid`` an atom-centred quadrature (Gauss-Chebyshev radial x Gauss-Legendre/uniform angular product grid, Becke
                 partitioning).  NOT PySCF's grid (Treutler/Lebedev tables are not reproducible from memory): grids are an
                 INPUT of the device path (the caller passes PySCF's ``mf.grids.coords / weights``); this one exists so
                 the oracle can be converged and pinned.

Pin: the reference's golden global B3LYP energy of water / STO-3G, ``-75.3091447400438`` with ``energy_elec`` =
``(-84.59485896172163, 37.93302591280513)`` (tests/test_driver.py:45-49), is reproduced by this restatement on a
converged grid (tests/test_golden.py::test_xc_restatement_reproduces_the_reference_b3lyp_energy).  That run is closed
shell: the spin-polarised branches (zeta != 0 of VWN, rho_a != rho_b of LYP/B88) follow the published formulas but are
unpinned by any golden of the reference.
"""
from __future__ import annotations

import math

import numpy as np

from . import gto_restatement as g
from . import pyscf_restatement as ps


# --------------------------------------------------------------------------------------------
# forward-mode automatic differentiation over NumPy arrays: value + NV partial derivatives
# --------------------------------------------------------------------------------------------
class Jet:
    __slots__ = ("v", "d")

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def variables(*arrays):
        n = len(arrays)
        out = []
        for i, a in enumerate(arrays):
            a = np.asarray(a, dtype=np.float64)
            d = np.zeros((n,) + a.shape)
            d[i] = 1.0
            out.append(Jet(a, d))
        return out

    @staticmethod
    def lift(x, like):
        return x if isinstance(x, Jet) else Jet(np.zeros_like(like.v) + x, np.zeros_like(like.d))

    def __add__(self, o):
        o = Jet.lift(o, self)
        return Jet(self.v + o.v, self.d + o.d)

    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, -self.d)

    def __sub__(self, o):
        o = Jet.lift(o, self)
        return Jet(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return Jet.lift(o, self) - self

    def __mul__(self, o):
        o = Jet.lift(o, self)
        return Jet(self.v * o.v, self.d * o.v + o.d * self.v)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Jet.lift(o, self)
        inv = 1.0 / o.v
        q = self.v * inv
        return Jet(q, (self.d - o.d * q) * inv)

    def __rtruediv__(self, o):
        return Jet.lift(o, self) / self

    def __pow__(self, p):  # real constant exponent
        vp1 = self.v ** (p - 1.0)
        return Jet(vp1 * self.v, self.d * (p * vp1))


def _f(fn, dfn):
    def apply(x):
        if isinstance(x, Jet):
            return Jet(fn(x.v), x.d * dfn(x.v))
        return fn(x)

    return apply


jlog = _f(np.log, lambda v: 1.0 / v)
jexp = _f(np.exp, np.exp)
jsqrt = _f(np.sqrt, lambda v: 0.5 / np.sqrt(v))
jatan = _f(np.arctan, lambda v: 1.0 / (1.0 + v * v))
jasinh = _f(np.arcsinh, lambda v: 1.0 / np.sqrt(1.0 + v * v))


# --------------------------------------------------------------------------------------------
# functionals (energy per unit volume); arguments are Jets or arrays
# --------------------------------------------------------------------------------------------
CX = 0.75 * (3.0 / math.pi) ** (1.0 / 3.0)


def slater_x(ra, rb):
    return -CX * 2.0 ** (1.0 / 3.0) * (ra ** (4.0 / 3.0) + rb ** (4.0 / 3.0))


def b88_x(ra, rb, saa, sbb, beta=0.0042):
    """Becke 1988 exchange INCLUDING its LDA part (libxc GGA_X_B88)."""
    out = slater_x(ra, rb)
    for r, s in ((ra, saa), (rb, sbb)):
        r43 = r ** (4.0 / 3.0)
        x = jsqrt(s) / r43
        out = out - beta * r43 * x * x / (1.0 + 6.0 * beta * x * jasinh(x))
    return out


def _vwn_aux(a, b, c, x0, rs):
    x = jsqrt(rs)
    big_x = x * x + b * x + c
    q = math.sqrt(4.0 * c - b * b)
    x0x = x0 * x0 + b * x0 + c
    at = jatan(q / (2.0 * x + b))
    return a * (jlog(x * x / big_x) + (2.0 * b / q) * at
                - (b * x0 / x0x) * (jlog((x - x0) * (x - x0) / big_x) + (2.0 * (b + 2.0 * x0) / q) * at))


def vwn_rpa_c(ra, rb):
    """libxc LDA_C_VWN_RPA: RPA parameter set, para / ferro interpolated with f(zeta)."""
    rho = ra + rb
    rs = (3.0 / (4.0 * math.pi)) ** (1.0 / 3.0) * rho ** (-1.0 / 3.0)
    z = (ra - rb) / rho
    fz = ((1.0 + z) ** (4.0 / 3.0) + (1.0 - z) ** (4.0 / 3.0) - 2.0) / (2.0 ** (4.0 / 3.0) - 2.0)
    ep = _vwn_aux(0.0310907, 13.0720, 42.7198, -0.409286, rs)
    ef = _vwn_aux(0.01554535, 20.1231, 101.578, -0.743294, rs)
    return rho * (ep * (1.0 - fz) + ef * fz)


def lyp_c(ra, rb, saa, sab, sbb, a=0.04918, b=0.132, c=0.2533, d=0.349):
    """Lee-Yang-Parr correlation, spin-polarised form of Miehlich, Savin, Stoll, Preuss (CPL 157, 200 (1989))."""
    rho = ra + rb
    rm13 = rho ** (-1.0 / 3.0)
    den = 1.0 + d * rm13
    omega = jexp(-c * rm13) / den * rho ** (-11.0 / 3.0)
    delta = c * rm13 + d * rm13 / den
    cf = 0.3 * (3.0 * math.pi * math.pi) ** (2.0 / 3.0)
    sig = saa + 2.0 * sab + sbb
    t1 = -a * 4.0 / den * ra * rb / rho
    inner = ra * rb * (2.0 ** (11.0 / 3.0) * cf * (ra ** (8.0 / 3.0) + rb ** (8.0 / 3.0))
                       + (47.0 / 18.0 - 7.0 * delta / 18.0) * sig
                       - (2.5 - delta / 18.0) * (saa + sbb)
                       - (delta - 11.0) / 9.0 * (ra / rho * saa + rb / rho * sbb))
    inner = inner - (2.0 / 3.0) * rho * rho * sig + ((2.0 / 3.0) * rho * rho - ra * ra) * sbb + ((2.0 / 3.0) * rho * rho - rb * rb) * saa
    return t1 - a * b * omega * inner


HYB = {"b3lyp": 0.2, "lda": 0.0, "hf": 1.0}


def xc_energy_density(name, ra, rb, saa, sab, sbb):
    if name == "b3lyp":
        return (0.08 * slater_x(ra, rb) + 0.72 * b88_x(ra, rb, saa, sbb) + 0.19 * vwn_rpa_c(ra, rb)
                + 0.81 * lyp_c(ra, rb, saa, sab, sbb))
    if name == "lda":  # Slater exchange + VWN(RPA) correlation (a pure LDA for tests of the LDA code path)
        return slater_x(ra, rb) + vwn_rpa_c(ra, rb) + 0.0 * (saa + sab + sbb)
    raise ValueError(name)


RHO_CUT = 1e-10  # grid points with a spin density below this carry no XC (numint's small-density screening)


def eval_xc(name, ra, rb, saa, sab, sbb):
    """(f, vrho [2], vsigma [3]) per grid point; points with rho_a or rho_b < RHO_CUT give zeros."""
    ok = (ra > RHO_CUT) & (rb > RHO_CUT)
    f = np.zeros_like(ra)
    vr = np.zeros((2,) + ra.shape)
    vs = np.zeros((3,) + ra.shape)
    if ok.any():
        jets = Jet.variables(ra[ok], rb[ok], np.maximum(saa[ok], 1e-40), sab[ok], np.maximum(sbb[ok], 1e-40))
        e = xc_energy_density(name, *jets)
        f[ok] = e.v
        vr[:, ok] = e.d[:2]
        vs[:, ok] = e.d[2:]
    return f, vr, vs


# --------------------------------------------------------------------------------------------
# AO values and gradients on grid points (libcint real-spherical conventions, as gto_restatement)
# --------------------------------------------------------------------------------------------
def eval_ao(shells, coords):
    """(4, ng, nao): values and d/dx, d/dy, d/dz   (pyscf.dft.numint.eval_ao(mol, coords, deriv=1))."""
    ng = coords.shape[0]
    nao = g.nao_sph(shells)
    out = np.zeros((4, ng, nao))
    off = 0
    for (ctr, l, exps, coefs) in shells:
        d = coords - ctr
        r2 = (d * d).sum(axis=1)
        rad = np.zeros(ng)
        drad = np.zeros(ng)  # d(rad)/d(r^2)
        for a, c in zip(exps, coefs):
            e = c * np.exp(-a * r2)
            rad += e
            drad += -a * e
        comps = g.cart_components(l)
        cart = np.zeros((4, ng, len(comps)))
        for k, (lx, ly, lz) in enumerate(comps):
            px, py, pz = d[:, 0] ** lx, d[:, 1] ** ly, d[:, 2] ** lz
            poly = px * py * pz
            cart[0, :, k] = poly * rad
            for ax, (ll, pw) in enumerate(((lx, (py * pz)), (ly, (px * pz)), (lz, (px * py)))):
                dpoly = ll * d[:, ax] ** (ll - 1) * pw if ll > 0 else 0.0
                cart[1 + ax, :, k] = dpoly * rad + poly * 2.0 * d[:, ax] * drad
        t = g.cart2sph(l)
        out[:, :, off : off + 2 * l + 1] = cart @ t.T
        off += 2 * l + 1
    return out


# --------------------------------------------------------------------------------------------
# pyscf.dft.numint.NumInt.nr_uks (LDA / GGA)
# --------------------------------------------------------------------------------------------
def nr_uks(name, ao, weights, dms):
    """Returns (nelec [2], exc, vxc [2, nao, nao]) for spin densities ``dms`` [2, nao, nao]."""
    phi = ao[0]
    rho, grad = [], []
    for s in range(2):
        t = phi @ dms[s]
        rho.append((t * phi).sum(axis=1))
        grad.append(np.array([2.0 * (t * ao[1 + k]).sum(axis=1) for k in range(3)]))
    saa = (grad[0] * grad[0]).sum(axis=0)
    sab = (grad[0] * grad[1]).sum(axis=0)
    sbb = (grad[1] * grad[1]).sum(axis=0)
    f, vr, vs = eval_xc(name, rho[0], rho[1], saa, sab, sbb)
    exc = float((weights * f).sum())
    vxc = []
    for s in range(2):
        o = 1 - s
        gvec = 2.0 * vs[0 if s == 0 else 2] * grad[s] + vs[1] * grad[o]  # [3, ng]
        m = phi * (0.5 * weights * vr[s])[:, None]
        for k in range(3):
            m = m + ao[1 + k] * (weights * gvec[k])[:, None]
        v = phi.T @ m
        vxc.append(v + v.T)
    nelec = [float((weights * rho[s]).sum()) for s in range(2)]
    return nelec, exc, np.array(vxc)


# --------------------------------------------------------------------------------------------
# atom-centred quadrature (Becke partitioning); NOT PySCF's grid - see the module docstring
# --------------------------------------------------------------------------------------------
BRAGG = {"H": 0.35 / g.BOHR, "C": 0.70 / g.BOHR, "N": 0.65 / g.BOHR, "O": 0.60 / g.BOHR}


def becke_grid(atoms, n_rad=80, n_theta=24, n_phi=48):
    """(coords [ng, 3], weights [ng])."""
    xs = np.cos(np.arange(1, n_rad + 1) * math.pi / (n_rad + 1))  # Gauss-Chebyshev, second kind
    wx = math.pi / (n_rad + 1) * np.sin(np.arange(1, n_rad + 1) * math.pi / (n_rad + 1)) ** 2
    ct, wt = np.polynomial.legendre.leggauss(n_theta)
    ph = (np.arange(n_phi) + 0.5) * 2.0 * math.pi / n_phi
    st = np.sqrt(1.0 - ct * ct)
    ang = np.array([[s * math.cos(p), s * math.sin(p), c] for c, s in zip(ct, st) for p in ph])
    wang = np.array([w * 2.0 * math.pi / n_phi for w in wt for _ in ph])
    centres = np.array([x for _, x in atoms])
    nat = len(atoms)
    rij = np.linalg.norm(centres[:, None] - centres[None], axis=2)
    all_c, all_w = [], []
    for ia, (sym, ctr) in enumerate(atoms):
        rm = BRAGG[sym]
        r = rm * (1.0 + xs) / (1.0 - xs)
        wr = wx / np.sqrt(1.0 - xs * xs) * 2.0 * rm / (1.0 - xs) ** 2 * r * r  # dr/dx, Chebyshev weight removed, r^2
        pts = ctr + (r[:, None, None] * ang[None]).reshape(-1, 3)
        w = (wr[:, None] * wang[None]).reshape(-1)
        # Becke cell function with the Bragg-Slater size adjustment
        dist = np.linalg.norm(pts[:, None] - centres[None], axis=2)  # [ng, nat]
        cell = np.ones((pts.shape[0], nat))
        for i in range(nat):
            for j in range(nat):
                if i == j:
                    continue
                mu = (dist[:, i] - dist[:, j]) / rij[i, j]
                chi = BRAGG[atoms[i][0]] / BRAGG[atoms[j][0]]
                u = (chi - 1.0) / (chi + 1.0)
                aij = min(0.5, max(-0.5, u / (u * u - 1.0)))
                nu = mu + aij * (1.0 - mu * mu)
                for _ in range(3):
                    nu = 1.5 * nu - 0.5 * nu ** 3
                cell[:, i] *= 0.5 * (1.0 - nu)
        all_c.append(pts)
        all_w.append(w * cell[:, ia] / cell.sum(axis=1))
    return np.vstack(all_c), np.concatenate(all_w)


# --------------------------------------------------------------------------------------------
# pyscf.dft.uks.UKS over density-fitted J/K: get_veff (tagged with ecoul / exc / vj / vk) and energy_elec
# --------------------------------------------------------------------------------------------
class DFUKS(ps.DFUHF):
    """Duck-typed ``pyscf.dft.UKS(...).density_fit()`` on explicit tensors and an explicit grid."""

    is_ks = True

    def __init__(self, ovlp, hcore, cderi, nelec, ao, weights, xc="b3lyp", **kw):
        super().__init__(ovlp, hcore, cderi, nelec, **kw)
        self.ao, self.weights, self.xc = ao, weights, xc
        self.n_xc_builds = 0

    def get_veff(self, mol=None, dm=None, dm_last=0, vhf_last=0, hermi=1):
        """pyscf/dft/uks.py:get_veff (no incremental build: DF objects set direct_scf = False)."""
        if dm is None:
            dm = self.make_rdm1()
        if isinstance(dm, np.ndarray) and dm.ndim == 2:
            dm = np.asarray((dm * 0.5, dm * 0.5))
        self.n_xc_builds += 1
        n, exc, vxc = nr_uks(self.xc, self.ao, self.weights, np.asarray(dm))
        hyb = HYB[self.xc]
        if abs(hyb) < 1e-10:
            vj = self.get_j(mol, dm, hermi)
            vj = vj[0] + vj[1]
            vxc = vxc + vj
            vk = None
        else:
            vj, vk = self.get_jk(mol, dm, hermi)
            vj = vj[0] + vj[1]
            vk = vk * hyb
            vxc = vxc + vj - vk
            exc -= (np.einsum("ij,ji", dm[0], vk[0]).real + np.einsum("ij,ji", dm[1], vk[1]).real) * 0.5
        ecoul = np.einsum("ij,ji", dm[0] + dm[1], vj).real * 0.5
        return ps.tag_array(vxc, ecoul=ecoul, exc=exc, vj=vj, vk=vk)

    def energy_elec(self, dm=None, h1e=None, vhf=None):
        """pyscf/dft/uks.py:energy_elec (= rks.energy_elec on the total density): e1 + ecoul + exc."""
        if dm is None:
            dm = self.make_rdm1()
        if h1e is None:
            h1e = self.get_hcore()
        if vhf is None or getattr(vhf, "ecoul", None) is None:
            vhf = self.get_veff(dm=dm)
        dmt = np.asarray(dm)
        dmt = dmt[0] + dmt[1] if dmt.ndim == 3 else dmt
        h1e = np.asarray(h1e)
        if h1e.ndim == 3:  # not PySCF: the patched spin-resolved core Hamiltonian goes through nbed's own energy_elec
            e1 = np.einsum("sij,sji->", h1e, np.asarray(dm)).real
        else:
            e1 = np.einsum("ij,ji->", h1e, dmt).real
        e2 = vhf.ecoul + vhf.exc
        self.scf_summary.update(e1=e1, coul=vhf.ecoul, exc=vhf.exc)
        return e1 + e2, e2


class DFRKS(ps.DFRHF):
    """Duck-typed ``pyscf.dft.RKS(...).density_fit()`` on explicit tensors and an explicit grid (the object type of the
    reference's ``tests/test_scf.py:19-40``; its drivers only build UKS objects, nbed/driver.py:289-313)."""

    is_ks = True

    def __init__(self, ovlp, hcore, cderi, nelec, ao, weights, xc="b3lyp", **kw):
        super().__init__(ovlp, hcore, cderi, nelec, **kw)
        self.ao, self.weights, self.xc = ao, weights, xc
        self.n_xc_builds = 0

    def get_veff(self, mol=None, dm=None, dm_last=0, vhf_last=0, hermi=1):
        """pyscf/dft/rks.py:get_veff: vxc + vj - hyb / 2 vk on the total density, tagged ecoul / exc / vj / vk.
        ``nr_rks`` is ``nr_uks`` on the spin-unpolarised pair (dm / 2, dm / 2): same energy, vxc of either spin."""
        if dm is None:
            dm = self.make_rdm1()
        dm = np.asarray(dm)
        self.n_xc_builds += 1
        _, exc, vxc2 = nr_uks(self.xc, self.ao, self.weights, np.asarray((dm * 0.5, dm * 0.5)))
        vxc = vxc2[0]
        hyb = HYB[self.xc]
        if abs(hyb) < 1e-10:
            vj = self.get_j(mol, dm, hermi)
            vxc = vxc + vj
            vk = None
        else:
            vj, vk = self.get_jk(mol, dm, hermi)
            vk = vk * hyb
            vxc = vxc + vj - vk * 0.5
            exc -= np.einsum("ij,ji", dm, vk).real * 0.5 * 0.5
        ecoul = np.einsum("ij,ji", dm, vj).real * 0.5
        return ps.tag_array(vxc, ecoul=ecoul, exc=exc, vj=vj, vk=vk)

    def energy_elec(self, dm=None, h1e=None, vhf=None):
        """pyscf/dft/rks.py:energy_elec: e1 + ecoul + exc."""
        if dm is None:
            dm = self.make_rdm1()
        if h1e is None:
            h1e = self.get_hcore()
        if vhf is None or getattr(vhf, "ecoul", None) is None:
            vhf = self.get_veff(dm=dm)
        e1 = np.einsum("ij,ji->", np.asarray(h1e), np.asarray(dm)).real
        e2 = vhf.ecoul + vhf.exc
        self.scf_summary.update(e1=e1, coul=vhf.ecoul, exc=vhf.exc)
        return e1 + e2, e2

