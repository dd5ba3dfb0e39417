"""Minimal Cartesian s/p Gaussian integral generator (McMurchie–Davidson).  TEST INFRASTRUCTURE.

libcint (the reference's integral source, reached through pyscf.gto) is absent from this image, so the
water/STO-3G system of the reference's tests (tests/conftest.py:28-36, tests/molecules/water.xyz) is
rebuilt here to pin the oracle on the reference's golden energies (tests/test_driver.py:56-57,76).
Conventions follow PySCF: Bohr = 0.52917721092 Angstrom, AO order per atom 1s,2s,2px,2py,2pz,
contracted functions normalised.  Only s and p shells are needed for STO-3G first-row atoms.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
from scipy.special import hyp1f1

BOHR = 0.52917721092

# STO-3G (EMSL / PySCF 'sto-3g'): (l, exponents, contraction coefficients)
_SP = [0.15432897, 0.53532814, 0.44463454]
STO3G = {
    "H": [(0, [3.42525091, 0.62391373, 0.16885540], _SP)],
    "O": [
        (0, [130.7093200, 23.8088610, 6.4436083], _SP),
        (0, [5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547]),
        (1, [5.0331513, 1.1695961, 0.3803890], [0.15591627, 0.60768372, 0.39195739]),
    ],
}
CHARGE = {"H": 1, "O": 8}


def parse_xyz(xyz: str):
    lines = [ln for ln in xyz.strip("\n").split("\n")]
    natm = int(lines[0].split()[0])
    atoms = []
    for ln in lines[2 : 2 + natm]:
        t = ln.split()
        atoms.append((t[0], np.array([float(v) for v in t[1:4]]) / BOHR))
    return atoms


def _norm_prim(a, lmn):
    l, m, n = lmn
    L = l + m + n
    df = lambda k: 1.0 if k <= 0 else float(np.prod(np.arange(k, 0, -2)))  # noqa: E731
    return math.sqrt(
        (2 * a / math.pi) ** 1.5 * (4 * a) ** L / (df(2 * l - 1) * df(2 * m - 1) * df(2 * n - 1))
    )


class BasisFunction:
    def __init__(self, center, lmn, exps, coefs):
        self.center = np.asarray(center, float)
        self.lmn = lmn
        self.exps = list(exps)
        self.coefs = np.array(coefs, float) * np.array([_norm_prim(a, lmn) for a in exps])
        # normalise the contraction
        s = 0.0
        for (a, ca), (b, cb) in itertools.product(zip(self.exps, self.coefs), repeat=2):
            s += ca * cb * _overlap_prim(a, lmn, self.center, b, lmn, self.center)
        self.coefs /= math.sqrt(s)


def build_basis(atoms, basis=STO3G):
    fns = []
    for sym, xyz in atoms:
        for l, exps, coefs in basis[sym]:
            if l == 0:
                fns.append(BasisFunction(xyz, (0, 0, 0), exps, coefs))
            else:
                for lmn in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
                    fns.append(BasisFunction(xyz, lmn, exps, coefs))
    return fns


def _E(i, j, t, Qx, a, b):
    """Hermite expansion coefficient E_t^{ij}."""
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return math.exp(-q * Qx * Qx)
    if j == 0:
        return (1 / (2 * p)) * _E(i - 1, j, t - 1, Qx, a, b) - (q * Qx / a) * _E(i - 1, j, t, Qx, a, b) + (
            t + 1
        ) * _E(i - 1, j, t + 1, Qx, a, b)
    return (1 / (2 * p)) * _E(i, j - 1, t - 1, Qx, a, b) + (q * Qx / b) * _E(i, j - 1, t, Qx, a, b) + (
        t + 1
    ) * _E(i, j - 1, t + 1, Qx, a, b)


def _overlap_prim(a, lmn1, A, b, lmn2, B):
    s = 1.0
    for k in range(3):
        s *= _E(lmn1[k], lmn2[k], 0, A[k] - B[k], a, b)
    return s * (math.pi / (a + b)) ** 1.5


def _kinetic_prim(a, lmn1, A, b, lmn2, B):
    l2, m2, n2 = lmn2
    t0 = b * (2 * (l2 + m2 + n2) + 3) * _overlap_prim(a, lmn1, A, b, lmn2, B)
    t1 = -2 * b * b * (
        _overlap_prim(a, lmn1, A, b, (l2 + 2, m2, n2), B)
        + _overlap_prim(a, lmn1, A, b, (l2, m2 + 2, n2), B)
        + _overlap_prim(a, lmn1, A, b, (l2, m2, n2 + 2), B)
    )
    t2 = -0.5 * (
        l2 * (l2 - 1) * _overlap_prim(a, lmn1, A, b, (l2 - 2, m2, n2), B)
        + m2 * (m2 - 1) * _overlap_prim(a, lmn1, A, b, (l2, m2 - 2, n2), B)
        + n2 * (n2 - 1) * _overlap_prim(a, lmn1, A, b, (l2, m2, n2 - 2), B)
    )
    return t0 + t1 + t2


def _boys(n, T):
    return hyp1f1(n + 0.5, n + 1.5, -T) / (2.0 * n + 1.0)


def _R(t, u, v, n, p, PC, RPC2):
    """Hermite Coulomb integral R^n_{tuv}."""
    if t == u == v == 0:
        return (-2 * p) ** n * _boys(n, p * RPC2)
    val = 0.0
    if t > 0:
        if t > 1:
            val += (t - 1) * _R(t - 2, u, v, n + 1, p, PC, RPC2)
        val += PC[0] * _R(t - 1, u, v, n + 1, p, PC, RPC2)
    elif u > 0:
        if u > 1:
            val += (u - 1) * _R(t, u - 2, v, n + 1, p, PC, RPC2)
        val += PC[1] * _R(t, u - 1, v, n + 1, p, PC, RPC2)
    else:
        if v > 1:
            val += (v - 1) * _R(t, u, v - 2, n + 1, p, PC, RPC2)
        val += PC[2] * _R(t, u, v - 1, n + 1, p, PC, RPC2)
    return val


def _nuclear_prim(a, lmn1, A, b, lmn2, B, C):
    p = a + b
    P = (a * A + b * B) / p
    PC = P - C
    RPC2 = float(PC @ PC)
    val = 0.0
    for t in range(lmn1[0] + lmn2[0] + 1):
        Ex = _E(lmn1[0], lmn2[0], t, A[0] - B[0], a, b)
        for u in range(lmn1[1] + lmn2[1] + 1):
            Ey = _E(lmn1[1], lmn2[1], u, A[1] - B[1], a, b)
            for v in range(lmn1[2] + lmn2[2] + 1):
                Ez = _E(lmn1[2], lmn2[2], v, A[2] - B[2], a, b)
                val += Ex * Ey * Ez * _R(t, u, v, 0, p, PC, RPC2)
    return val * 2 * math.pi / p


def _hermite_pair(a, lmn1, A, b, lmn2, B):
    """List of (t,u,v, E_t E_u E_v) for a primitive pair."""
    out = []
    for t in range(lmn1[0] + lmn2[0] + 1):
        Ex = _E(lmn1[0], lmn2[0], t, A[0] - B[0], a, b)
        for u in range(lmn1[1] + lmn2[1] + 1):
            Ey = _E(lmn1[1], lmn2[1], u, A[1] - B[1], a, b)
            for v in range(lmn1[2] + lmn2[2] + 1):
                Ez = _E(lmn1[2], lmn2[2], v, A[2] - B[2], a, b)
                out.append((t, u, v, Ex * Ey * Ez))
    return out


def _eri_contracted(f1, f2, f3, f4, pair_cache):
    def pairs(fa, fb):
        key = (id(fa), id(fb))
        if key not in pair_cache:
            lst = []
            for a, ca in zip(fa.exps, fa.coefs):
                for b, cb in zip(fb.exps, fb.coefs):
                    p = a + b
                    P = (a * fa.center + b * fb.center) / p
                    lst.append((p, P, ca * cb, _hermite_pair(a, fa.lmn, fa.center, b, fb.lmn, fb.center)))
            pair_cache[key] = lst
        return pair_cache[key]

    val = 0.0
    for p, P, c12, h12 in pairs(f1, f2):
        for q, Q, c34, h34 in pairs(f3, f4):
            alpha = p * q / (p + q)
            PQ = P - Q
            RPQ2 = float(PQ @ PQ)
            s = 0.0
            for t, u, v, e12 in h12:
                if e12 == 0.0:
                    continue
                for tau, nu, phi, e34 in h34:
                    if e34 == 0.0:
                        continue
                    s += e12 * e34 * (-1) ** (tau + nu + phi) * _R(t + tau, u + nu, v + phi, 0, alpha, PQ, RPQ2)
            val += c12 * c34 * s * 2 * math.pi**2.5 / (p * q * math.sqrt(p + q))
    return val


def integrals(atoms, basis=STO3G):
    """Returns dict(S, T, V, eri (chemist (ij|kl)), e_nuc, nelectron)."""
    fns = build_basis(atoms, basis)
    n = len(fns)
    S = np.zeros((n, n))
    T = np.zeros((n, n))
    V = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            fi, fj = fns[i], fns[j]
            s = t = v = 0.0
            for a, ca in zip(fi.exps, fi.coefs):
                for b, cb in zip(fj.exps, fj.coefs):
                    s += ca * cb * _overlap_prim(a, fi.lmn, fi.center, b, fj.lmn, fj.center)
                    t += ca * cb * _kinetic_prim(a, fi.lmn, fi.center, b, fj.lmn, fj.center)
                    for sym, C in atoms:
                        v -= CHARGE[sym] * ca * cb * _nuclear_prim(a, fi.lmn, fi.center, b, fj.lmn, fj.center, C)
            S[i, j] = S[j, i] = s
            T[i, j] = T[j, i] = t
            V[i, j] = V[j, i] = v
    eri = np.zeros((n, n, n, n))
    cache = {}
    for i in range(n):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            for k in range(n):
                for l in range(k + 1):
                    kl = k * (k + 1) // 2 + l
                    if ij < kl:
                        continue
                    val = _eri_contracted(fns[i], fns[j], fns[k], fns[l], cache)
                    for a, b in ((i, j), (j, i)):
                        for c, d in ((k, l), (l, k)):
                            eri[a, b, c, d] = val
                            eri[c, d, a, b] = val
    e_nuc = 0.0
    for (s1, r1), (s2, r2) in itertools.combinations(atoms, 2):
        e_nuc += CHARGE[s1] * CHARGE[s2] / np.linalg.norm(r1 - r2)
    return {"S": S, "T": T, "V": V, "eri": eri, "e_nuc": e_nuc, "nelectron": sum(CHARGE[s] for s, _ in atoms)}
