"""Fock-space matrix of a second-quantised Hamiltonian and its Jordan-Wigner Pauli terms.  TEST INFRASTRUCTURE.

Stands in for ``openfermion`` (absent here) in the reference's builder tests
(tests/test_builder.py:55-120): ``get_sparse_operator(jordan_wigner(InteractionOperator(const, h1, h2)))``
is, by construction, the matrix  const + sum h1[p,q] a+_p a_q + sum h2[p,q,r,s] a+_p a+_q a_r a_s  on the
2^n-dimensional Fock space, which is what ``sparse_hamiltonian`` builds directly.
``pauli_terms`` enumerates the Jordan-Wigner Pauli strings with their coefficients so that the
"identical Pauli-term set" parity criterion of the north star can be checked without openfermion.
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg


def _annihilators(n):
    """JW annihilation operators, qubit 0 = most significant (openfermion convention)."""
    I2 = sp.identity(2, format="csr")
    Z = sp.csr_matrix(np.array([[1.0, 0.0], [0.0, -1.0]]))
    A = sp.csr_matrix(np.array([[0.0, 1.0], [0.0, 0.0]]))  # |0><1|
    ops = []
    for p in range(n):
        m = sp.identity(1, format="csr")
        for q in range(n):
            m = sp.kron(m, Z if q < p else (A if q == p else I2), format="csr")
        ops.append(m)
    return ops


def sparse_hamiltonian(const, h1, h2):
    n = h1.shape[0]
    a = _annihilators(n)
    ad = [x.T.tocsr() for x in a]
    dim = 2**n
    H = const * sp.identity(dim, format="csr")
    for p, q in itertools.product(range(n), repeat=2):
        if h1[p, q] != 0.0:
            H = H + h1[p, q] * (ad[p] @ a[q])
    # pair operators
    rs_ops = {}
    for r, s in itertools.product(range(n), repeat=2):
        if r != s:
            rs_ops[(r, s)] = a[r] @ a[s]
    for p, q in itertools.product(range(n), repeat=2):
        if p == q:
            continue
        acc = None
        for (r, s), op in rs_ops.items():
            v = h2[p, q, r, s]
            if v != 0.0:
                acc = v * op if acc is None else acc + v * op
        if acc is not None:
            H = H + (ad[p] @ ad[q]) @ acc
    return H.tocsr()


def ground_energies(const, h1, h2, k=1):
    H = sparse_hamiltonian(const, h1, h2)
    w = scipy.sparse.linalg.eigsh(H, k=k, which="SA", return_eigenvectors=False, tol=1e-12)
    return np.sort(w)


# ---- Jordan-Wigner Pauli expansion -----------------------------------------------------------
# a_p   = (prod_{q<p} Z_q) (X_p + iY_p)/2 ;  a+_p = (prod_{q<p} Z_q) (X_p - iY_p)/2
_MUL = {  # single-qubit Pauli products: (a, b) -> (phase, c)
    ("I", "I"): (1, "I"), ("I", "X"): (1, "X"), ("I", "Y"): (1, "Y"), ("I", "Z"): (1, "Z"),
    ("X", "I"): (1, "X"), ("X", "X"): (1, "I"), ("X", "Y"): (1j, "Z"), ("X", "Z"): (-1j, "Y"),
    ("Y", "I"): (1, "Y"), ("Y", "X"): (-1j, "Z"), ("Y", "Y"): (1, "I"), ("Y", "Z"): (1j, "X"),
    ("Z", "I"): (1, "Z"), ("Z", "X"): (1j, "Y"), ("Z", "Y"): (-1j, "X"), ("Z", "Z"): (1, "I"),
}


def _pmul(t1, t2):
    """Multiply two Pauli-string dicts {string tuple: coeff}."""
    out = {}
    for s1, c1 in t1.items():
        for s2, c2 in t2.items():
            ph = c1 * c2
            s = []
            for x, y in zip(s1, s2):
                f, z = _MUL[(x, y)]
                ph *= f
                s.append(z)
            s = tuple(s)
            out[s] = out.get(s, 0) + ph
    return out


def _ladder(n, p, dagger):
    base = ["Z"] * p + ["I"] * (n - p)
    sx = list(base)
    sx[p] = "X"
    sy = list(base)
    sy[p] = "Y"
    return {tuple(sx): 0.5, tuple(sy): (-0.5j if dagger else 0.5j)}


def pauli_terms(const, h1, h2, tol=1e-12):
    """{pauli string (e.g. 'XZZY...'): real coefficient} of the JW-transformed Hamiltonian."""
    n = h1.shape[0]
    lad = [(_ladder(n, p, False), _ladder(n, p, True)) for p in range(n)]
    out = {("I",) * n: complex(const)}

    def add(terms, w):
        for s, c in terms.items():
            out[s] = out.get(s, 0) + w * c

    one_cache = {}
    for p, q in itertools.product(range(n), repeat=2):
        if h1[p, q] != 0.0:
            one_cache[(p, q)] = _pmul(lad[p][1], lad[q][0])
            add(one_cache[(p, q)], h1[p, q])
    dd = {}
    aa = {}
    nz = np.argwhere(h2 != 0.0)
    for p, q, r, s in nz:
        if p == q or r == s:
            continue
        if (p, q) not in dd:
            dd[(p, q)] = _pmul(lad[p][1], lad[q][1])
        if (r, s) not in aa:
            aa[(r, s)] = _pmul(lad[r][0], lad[s][0])
        add(_pmul(dd[(p, q)], aa[(r, s)]), h2[p, q, r, s])
    return {"".join(s): c.real for s, c in out.items() if abs(c) > tol}
