"""General-l Gaussian integrals in libcint / PySCF conventions (McMurchie-Davidson).  TEST INFRASTRUCTURE.

The reference gets its integrals from libcint through ``pyscf.gto`` / ``pyscf.df`` (``mf.density_fit()`` ->
``df.incore.cholesky_eri`` -> ``aux_e2(mol, auxmol, 'int3c2e', aosym='s2ij')`` and ``auxmol.intor('int2c2e')``); neither
library is in this image.  This module restates what the device generator of SURVEY.md section 8(f) rank 4 must
reproduce, independently of it (table-driven NumPy here, per-thread recursions in CUDA there):

* ``make_env``           pyscf.gto.mole.make_env / make_bas_env: the (atm, bas, env) arrays libcint consumes, with
                         PySCF's primitive normalisation (``gto_norm``) and contraction normalisation
                         (``_nomalize_contracted_ao``) folded into the coefficients stored in ``env``;
* ``cart2sph``           libcint's real-spherical transformation (``c2s``): s and p carry the constant factors
                         ``CINTcommon_fac_sp``, l >= 2 the unit-normalised real solid harmonics, m = -l..l, Cartesian
                         components in libcint order (xx, xy, xz, yy, yz, zz);
* ``int3c2e_sph`` / ``int2c2e_sph`` / ``int1e_*``    the integrals themselves;
* ``cholesky_eri``       pyscf.df.incore.cholesky_eri: ``cderi = L^-1 (P|mu nu)``, ``(P|Q) = L L^T``, packed lower rows.

s/p shells are pinned on the reference's golden water/STO-3G energies through oracle/gaussian_integrals.py (the same
recursions); higher l has no golden in the reference's tests and is validated by rotational invariance and by the
full-rank identity against exact four-centre integrals (tests/test_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg
from scipy.special import hyp1f1

BOHR = 0.52917721092
CHARGE = {"H": 1, "C": 6, "N": 7, "O": 8}

# libcint slots (cint.h)
ATM_SLOTS, BAS_SLOTS = 6, 8
CHARGE_OF, PTR_COORD = 0, 1
ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, KAPPA_OF, PTR_EXP, PTR_COEFF = 0, 1, 2, 3, 4, 5, 6
PTR_ENV_START = 20

# cc-pVDZ (Dunning 1989; EMSL / PySCF 'cc-pvdz'), segmented here: one shell per contracted function
_O_S_EXP = [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013]
CCPVDZ = {
    "H": [
        (0, [13.01, 1.962, 0.4446], [0.019685, 0.137977, 0.478148]),
        (0, [0.122], [1.0]),
        (1, [0.727], [1.0]),
    ],
    "O": [
        (0, _O_S_EXP, [0.000710, 0.005470, 0.027837, 0.104800, 0.283062, 0.448719, 0.270952, 0.015458]),
        (0, _O_S_EXP, [-0.000160, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368]),
        (0, [0.3023], [1.0]),
        (1, [17.70, 3.854, 1.046], [0.043018, 0.228913, 0.508728]),
        (1, [0.2753], [1.0]),
        (2, [1.185], [1.0]),
    ],
}


def even_tempered_aux(basis, lmax_aux=None, beta=2.0):
    """Even-tempered auxiliary basis in the spirit of pyscf.df.addons.aug_etb: per atom and per l <= 2 l_max(AO) an
    uncontracted geometric series covering [2 a_min, 2 a_max] of the AO exponents that can form that l by product."""
    aux = {}
    for sym, shells in basis.items():
        lmax = max(l for l, _, _ in shells)
        emin = {l: min(min(e) for ll, e, _ in shells if ll == l) for l in range(lmax + 1)}
        emax = {l: max(max(e) for ll, e, _ in shells if ll == l) for l in range(lmax + 1)}
        out = []
        top = 2 * lmax if lmax_aux is None else min(2 * lmax, lmax_aux)
        for L in range(top + 1):
            lo = min(emin[a] + emin[b] for a in range(lmax + 1) for b in range(lmax + 1) if abs(a - b) <= L <= a + b)
            hi = max(emax[a] + emax[b] for a in range(lmax + 1) for b in range(lmax + 1) if abs(a - b) <= L <= a + b)
            hi = min(hi, lo * beta ** 11)
            k = max(1, int(math.ceil(math.log(hi / lo) / math.log(beta))) + 1)
            for i in range(k):
                out.append((L, [lo * beta ** i], [1.0]))
        aux[sym] = out
    return aux


def parse_xyz(xyz: str):
    lines = xyz.strip("\n").split("\n")
    natm = int(lines[0].split()[0])
    atoms = []
    for ln in lines[2 : 2 + natm]:
        t = ln.split()
        atoms.append((t[0], np.array([float(v) for v in t[1:4]]) / BOHR))
    return atoms


# --------------------------------------------------------------------------------------------
# pyscf.gto.mole: gto_norm, _nomalize_contracted_ao, make_env
# --------------------------------------------------------------------------------------------
def gaussian_int(n, alpha):
    """int_0^inf r^n exp(-alpha r^2) dr   (pyscf.gto.mole.gaussian_int)"""
    n1 = (n + 1) * 0.5
    return math.gamma(n1) / (2.0 * alpha ** n1)


def gto_norm(l, expnt):
    """1 / sqrt(int r^(2l+2) exp(-2 a r^2) dr)   (pyscf.gto.mole.gto_norm)"""
    return 1.0 / math.sqrt(gaussian_int(l * 2 + 2, 2 * expnt))


def normalized_coefficients(l, exps, coefs):
    """Coefficients as PySCF stores them in ``mol._env``: primitive norms folded in, contraction normalised."""
    es = np.asarray(exps, float)
    cs = np.asarray(coefs, float) * np.array([gto_norm(l, e) for e in es])
    ee = es[:, None] + es[None, :]
    ee = np.vectorize(lambda a: gaussian_int(l * 2 + 2, a))(ee)
    s1 = 1.0 / math.sqrt(float(cs @ ee @ cs))
    return cs * s1


def make_env(atoms, basis, pre_atm=None, pre_env=None):
    """(atm, bas, env) in libcint layout for ``atoms`` = [(symbol, xyz_bohr)] and ``basis`` = {symbol: [(l, exps, coefs)]}."""
    env = list(np.zeros(PTR_ENV_START)) if pre_env is None else list(pre_env)
    atm, bas = [], []
    for ia, (sym, xyz) in enumerate(atoms):
        ptr = len(env)
        env.extend([float(x) for x in xyz])
        env.append(0.0)  # zeta slot
        atm.append([CHARGE[sym], ptr, 1, ptr + 3, 0, 0])
        for l, exps, coefs in basis[sym]:
            pe = len(env)
            env.extend([float(e) for e in exps])
            pc = len(env)
            env.extend([float(c) for c in normalized_coefficients(l, exps, coefs)])
            bas.append([ia, l, len(exps), 1, 0, pe, pc, 0])
    return np.array(atm, dtype=np.int32).reshape(-1, ATM_SLOTS), np.array(bas, dtype=np.int32).reshape(-1, BAS_SLOTS), np.array(env)


def conc_env(atoms, basis_ao, basis_aux):
    """``mol + auxmol`` concatenated the way pyscf.df.incore.aux_e2 does (gto.mole.conc_mol): AO shells first.
    Returns (atm, bas, env, nbas_ao)."""
    atm1, bas1, env1 = make_env(atoms, basis_ao)
    atm2, bas2, env2 = make_env(atoms, basis_aux, pre_env=env1)
    bas2 = bas2.copy()
    bas2[:, ATOM_OF] += len(atm1)
    return np.vstack([atm1, atm2]), np.vstack([bas1, bas2]), env2, len(bas1)


def shells_from_env(atm, bas, env, lo=0, hi=None):
    out = []
    for b in bas[lo:hi]:
        c = env[atm[b[ATOM_OF], PTR_COORD] : atm[b[ATOM_OF], PTR_COORD] + 3]
        out.append((np.array(c), int(b[ANG_OF]), env[b[PTR_EXP] : b[PTR_EXP] + b[NPRIM_OF]], env[b[PTR_COEFF] : b[PTR_COEFF] + b[NPRIM_OF]]))
    return out


def nao_sph(shells):
    return sum(2 * l + 1 for _, l, _, _ in shells)


# --------------------------------------------------------------------------------------------
# libcint c2s: Cartesian components (xx, xy, xz, yy, yz, zz order) -> real spherical, m = -l..l
# --------------------------------------------------------------------------------------------
def cart_components(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def cart2sph(l):
    """(2l+1, ncart) matrix.  l = 0, 1: libcint's CINTcommon_fac_sp constants (p keeps the order x, y, z);
    l >= 2: unit-normalised real solid harmonics (Helgaker, Jorgensen, Olsen eq. 6.4.47-6.4.50)."""
    comps = cart_components(l)
    if l == 0:
        return np.array([[0.282094791773878143]])
    if l == 1:
        return 0.488602511902919921 * np.eye(3)
    idx = {c: i for i, c in enumerate(comps)}
    out = np.zeros((2 * l + 1, len(comps)))
    for m in range(-l, l + 1):
        am = abs(m)
        nlm = (1.0 / (2 ** am * math.factorial(l))) * math.sqrt(
            2.0 * math.factorial(l + am) * math.factorial(l - am) / (2.0 if m == 0 else 1.0))
        vm2 = 1 if m < 0 else 0  # 2 * v_m
        for t in range((l - am) // 2 + 1):
            for u in range(t + 1):
                v2 = vm2
                while v2 <= am:  # 2 v runs over vm2, vm2 + 2, ... <= |m|
                    c = ((-1) ** (t + (v2 - vm2) // 2) * 0.25 ** t * math.comb(l, t) * math.comb(l - t, am + t)
                         * math.comb(t, u) * math.comb(am, v2))
                    ex = 2 * t + am - 2 * u - v2
                    ey = 2 * u + v2
                    ez = l - 2 * t - am
                    out[m + l, idx[(ex, ey, ez)]] += nlm * c
                    v2 += 2
        out[m + l] *= math.sqrt((2 * l + 1) / (4.0 * math.pi))
    return out


# --------------------------------------------------------------------------------------------
# McMurchie-Davidson tables
# --------------------------------------------------------------------------------------------
def hermite_E(la, lb, a, b, ab):
    """E[i, j, t] for one Cartesian direction: x_A^i x_B^j exp(-a x_A^2 - b x_B^2) = sum_t E_t^{ij} Lambda_t."""
    p = a + b
    mu = a * b / p
    e = np.zeros((la + 1, lb + 1, la + lb + 2))
    e[0, 0, 0] = math.exp(-mu * ab * ab)
    xpa, xpb = -b / p * ab, a / p * ab
    for i in range(la + 1):
        for j in range(lb + 1):
            if i == 0 and j == 0:
                continue
            for t in range(i + j + 1):
                if j == 0:
                    v = xpa * e[i - 1, j, t] + (t + 1) * e[i - 1, j, t + 1]
                    if t > 0:
                        v += e[i - 1, j, t - 1] / (2 * p)
                else:
                    v = xpb * e[i, j - 1, t] + (t + 1) * e[i, j - 1, t + 1]
                    if t > 0:
                        v += e[i, j - 1, t - 1] / (2 * p)
                e[i, j, t] = v
    return e[:, :, : la + lb + 1]


def hermite_E1(l, g):
    """Single Gaussian x^i exp(-g x^2) = sum_t E_t^i Lambda_t (the b -> 0 limit of hermite_E)."""
    return hermite_E(l, 0, g, 0.0, 0.0)[:, 0, :]


def boys(nmax, t):
    n = np.arange(nmax + 1)
    return hyp1f1(n + 0.5, n + 1.5, -t) / (2.0 * n + 1.0)


def hermite_R(L, alpha, pq):
    """R[t, u, v] (t + u + v <= L) of the Coulomb operator, exponent alpha, separation pq."""
    r2 = float(pq @ pq)
    f = boys(L, alpha * r2)
    rn = np.zeros((L + 1, L + 1, L + 1, L + 1))  # [n, t, u, v]
    for n in range(L + 1):
        rn[n, 0, 0, 0] = (-2.0 * alpha) ** n * f[n]
    for tot in range(1, L + 1):
        for t in range(tot + 1):
            for u in range(tot - t + 1):
                v = tot - t - u
                for n in range(L - tot + 1):
                    if t > 0:
                        val = pq[0] * rn[n + 1, t - 1, u, v]
                        if t > 1:
                            val += (t - 1) * rn[n + 1, t - 2, u, v]
                    elif u > 0:
                        val = pq[1] * rn[n + 1, t, u - 1, v]
                        if u > 1:
                            val += (u - 1) * rn[n + 1, t, u - 2, v]
                    else:
                        val = pq[2] * rn[n + 1, t, u, v - 1]
                        if v > 1:
                            val += (v - 1) * rn[n + 1, t, u, v - 2]
                    rn[n, t, u, v] = val
    return rn[0]


def _pair_hermite(sa, sb):
    """For a shell pair: list over primitive pairs of (p, P, coef, H[ca, cb, t, u, v])."""
    (A, la, ea, ca), (B, lb, eb, cb) = sa, sb
    compa, compb = cart_components(la), cart_components(lb)
    out = []
    ab = A - B
    for a, wa in zip(ea, ca):
        for b, wb in zip(eb, cb):
            p = a + b
            P = (a * A + b * B) / p
            ex, ey, ez = (hermite_E(la, lb, a, b, ab[k]) for k in range(3))
            L = la + lb
            h = np.zeros((len(compa), len(compb), L + 1, L + 1, L + 1))
            for i, (ax, ay, az) in enumerate(compa):
                for j, (bx, by, bz) in enumerate(compb):
                    h[i, j] = np.einsum("t,u,v->tuv", ex[ax, bx], ey[ay, by], ez[az, bz])
            out.append((p, P, wa * wb, h))
    return out


def _single_hermite(sc):
    (C, lc, ec, cc) = sc
    comp = cart_components(lc)
    out = []
    for g, w in zip(ec, cc):
        e1 = hermite_E1(lc, g)
        h = np.zeros((len(comp), lc + 1, lc + 1, lc + 1))
        for i, (cx, cy, cz) in enumerate(comp):
            h[i] = np.einsum("t,u,v->tuv", e1[cx], e1[cy], e1[cz])
        sign = np.fromfunction(lambda t, u, v: (-1.0) ** (t + u + v), h.shape[1:])
        out.append((g, C, w, h * sign))
    return out


def _coulomb_block(pairs, singles, lab, lc):
    """sum over primitives of  pref * sum_{tuv, t'u'v'} H_ab[tuv] Hc[t'u'v'] R[t+t', u+u', v+v']  -> [ca, cb, cc]."""
    L = lab + lc
    acc = None
    for p, P, wab, hab in pairs:
        for g, C, wc, hc in singles:
            alpha = p * g / (p + g)
            r = hermite_R(L, alpha, P - C)
            # shifted view: S[t, u, v, t', u', v'] = R[t + t', u + u', v + v']
            sh = np.zeros((lab + 1,) * 3 + (lc + 1,) * 3)
            for t in range(lab + 1):
                for u in range(lab + 1):
                    for v in range(lab + 1):
                        if t + u + v > lab:
                            continue
                        sh[t, u, v] = r[t : t + lc + 1, u : u + lc + 1, v : v + lc + 1]
            pref = wab * wc * 2.0 * math.pi ** 2.5 / (p * g * math.sqrt(p + g))
            term = pref * np.einsum("abtuv,tuvxyz,cxyz->abc", hab, sh, hc, optimize=True)
            acc = term if acc is None else acc + term
    return acc


def int3c2e_sph(ao_shells, aux_shells):
    """(naux, nao, nao) real-spherical three-centre integrals (P | mu nu)."""
    nao, naux = nao_sph(ao_shells), nao_sph(aux_shells)
    out = np.zeros((naux, nao, nao))
    singles = [_single_hermite(s) for s in aux_shells]
    tc = [cart2sph(s[1]) for s in aux_shells]
    off_a = np.cumsum([0] + [2 * s[1] + 1 for s in ao_shells])
    off_c = np.cumsum([0] + [2 * s[1] + 1 for s in aux_shells])
    for i, sa in enumerate(ao_shells):
        ta = cart2sph(sa[1])
        for j in range(i + 1):
            sb = ao_shells[j]
            tb = cart2sph(sb[1])
            pairs = _pair_hermite(sa, sb)
            for k, sc in enumerate(aux_shells):
                cart = _coulomb_block(pairs, singles[k], sa[1] + sb[1], sc[1])
                sph = np.einsum("ia,jb,kc,abc->kij", ta, tb, tc[k], cart, optimize=True)
                out[off_c[k] : off_c[k + 1], off_a[i] : off_a[i + 1], off_a[j] : off_a[j + 1]] = sph
                out[off_c[k] : off_c[k + 1], off_a[j] : off_a[j + 1], off_a[i] : off_a[i + 1]] = sph.transpose(0, 2, 1)
    return out


def int2c2e_sph(aux_shells):
    """(naux, naux) two-centre Coulomb metric (P | Q)."""
    naux = nao_sph(aux_shells)
    out = np.zeros((naux, naux))
    off = np.cumsum([0] + [2 * s[1] + 1 for s in aux_shells])
    one = (np.zeros(3), 0, [0.0], [1.0])  # the constant function 1 as the partner of a one-centre "pair"
    singles = [_single_hermite(s) for s in aux_shells]
    for i, sa in enumerate(aux_shells):
        ta = cart2sph(sa[1])
        # pair (sa, 1): exponent b = 0 keeps P = A and E^{i0}
        pairs = _pair_hermite(sa, (sa[0], 0, [0.0], [1.0]))
        for k in range(i + 1):
            sc = aux_shells[k]
            cart = _coulomb_block(pairs, singles[k], sa[1], sc[1])[:, 0, :]
            sph = ta @ cart @ cart2sph(sc[1]).T
            out[off[i] : off[i + 1], off[k] : off[k + 1]] = sph
            out[off[k] : off[k + 1], off[i] : off[i + 1]] = sph.T
    del one
    return out


# --------------------------------------------------------------------------------------------
# one-electron integrals (overlap, kinetic, nuclear attraction) in the spherical basis
# --------------------------------------------------------------------------------------------
def int1e_sph(ao_shells, atoms):
    """Returns (S, T, V) in the real-spherical AO basis; ``atoms`` = [(symbol, xyz_bohr)]."""
    nao = nao_sph(ao_shells)
    S, T, V = np.zeros((nao, nao)), np.zeros((nao, nao)), np.zeros((nao, nao))
    off = np.cumsum([0] + [2 * s[1] + 1 for s in ao_shells])
    for i, (A, la, ea, ca) in enumerate(ao_shells):
        ta = cart2sph(la)
        compa = cart_components(la)
        for j in range(i + 1):
            (B, lb, eb, cb) = ao_shells[j]
            tb = cart2sph(lb)
            compb = cart_components(lb)
            s = np.zeros((len(compa), len(compb)))
            t = np.zeros_like(s)
            v = np.zeros_like(s)
            ab = A - B
            for a, wa in zip(ea, ca):
                for b, wb in zip(eb, cb):
                    p = a + b
                    P = (a * A + b * B) / p
                    e = [hermite_E(la, lb + 2, a, b, ab[k]) for k in range(3)]
                    s1 = [e[k][:, :, 0] * math.sqrt(math.pi / p) for k in range(3)]  # 1-D overlaps [i, j]
                    # 1-D kinetic: -1/2 <i| d^2/dx^2 |j> = -2 b^2 S(i, j+2) + b (2j + 1) S(i, j) - j (j-1)/2 S(i, j-2)
                    k1 = []
                    for k in range(3):
                        kk = np.zeros((la + 1, lb + 1))
                        for jj in range(lb + 1):
                            kk[:, jj] = -2 * b * b * s1[k][:, jj + 2] + b * (2 * jj + 1) * s1[k][:, jj]
                            if jj >= 2:
                                kk[:, jj] -= 0.5 * jj * (jj - 1) * s1[k][:, jj - 2]
                        k1.append(kk)
                    rtabs = [(CHARGE[sym], hermite_R(la + lb, p, P - C)) for sym, C in atoms]
                    for x, (ax, ay, az) in enumerate(compa):
                        for y, (bx, by, bz) in enumerate(compb):
                            sx, sy, sz = s1[0][ax, bx], s1[1][ay, by], s1[2][az, bz]
                            s[x, y] += wa * wb * sx * sy * sz
                            t[x, y] += wa * wb * (k1[0][ax, bx] * sy * sz + sx * k1[1][ay, by] * sz + sx * sy * k1[2][az, bz])
                            h = np.einsum("t,u,v->tuv", e[0][ax, bx, : la + lb + 1], e[1][ay, by, : la + lb + 1], e[2][az, bz, : la + lb + 1])
                            for z, r in rtabs:
                                v[x, y] -= wa * wb * z * 2 * math.pi / p * float((h * r).sum())
            for m, c in ((S, s), (T, t), (V, v)):
                blk = ta @ c @ tb.T
                m[off[i] : off[i + 1], off[j] : off[j + 1]] = blk
                m[off[j] : off[j + 1], off[i] : off[i + 1]] = blk.T
    return S, T, V


def energy_nuc(atoms):
    e = 0.0
    for i in range(len(atoms)):
        for j in range(i):
            e += CHARGE[atoms[i][0]] * CHARGE[atoms[j][0]] / np.linalg.norm(atoms[i][1] - atoms[j][1])
    return e


# --------------------------------------------------------------------------------------------
# pyscf.df.incore.cholesky_eri
# --------------------------------------------------------------------------------------------
def cholesky_eri(j3c, j2c):
    """``cderi [naux, nao (nao + 1) / 2]`` = L^-1 (P | mu >= nu), (P|Q) = L L^T   (pyscf/df/incore.py:cholesky_eri)."""
    nao = j3c.shape[-1]
    il = np.tril_indices(nao)
    low = scipy.linalg.cholesky(j2c, lower=True)
    return scipy.linalg.solve_triangular(low, j3c[:, il[0], il[1]], lower=True, overwrite_b=False, check_finite=False)
