"""NumPy FP64 restatement of the PySCF 2.9.0 routines the Nbed hot path reaches.  TEST INFRASTRUCTURE.

PySCF (pinned ``pyscf==2.9.0``, /root/reference/uv.lock:2571-2572) is an un-vendored dependency that is
absent from this image, so its published algorithms are restated here and anchored on the reference's
own call sites:

  get_veff / get_jk       nbed/scf/huzinaga_scf.py:55,156 ; nbed/scf/embedded_hcore_funcs.py:34
  lib.diis.DIIS           nbed/scf/huzinaga_scf.py:130,164
  get_occ / make_rdm1     nbed/scf/huzinaga_scf.py:147-148,170-174
  scf.hf.kernel + CDIIS   nbed/driver.py:533  (mu-shift path)
  ao2mo.kernel / restore  nbed/ham_builder.py:128-141

The density-fitted route (``mf.density_fit()``: pyscf/df/df_jk.py:get_jk, pyscf/df/df_ao2mo.py) is the one
restated, because the north star mandates device-resident 3-centre tensors; exact-ERI parity for tiny
systems is obtained by feeding full-rank Cholesky vectors of the exact ERI (``cholesky_eri_exact``).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


# --------------------------------------------------------------------------------------------
# packed-lower helpers (pyscf.lib.pack_tril / unpack_tril; row-major lower triangle, i >= j)
# --------------------------------------------------------------------------------------------
def pack_tril(a: np.ndarray) -> np.ndarray:
    n = a.shape[-1]
    il = np.tril_indices(n)
    return np.ascontiguousarray(a[..., il[0], il[1]])


def unpack_tril(p: np.ndarray, n: int | None = None) -> np.ndarray:
    npair = p.shape[-1]
    if n is None:
        n = int((np.sqrt(8 * npair + 1) - 1) // 2)
    if p.ndim >= 2 and p.size >= 1 << 16:  # large blocks: C helper (what pyscf.lib.unpack_tril is), if built
        from . import c_binding

        out = c_binding.unpack_tril(p, n)
        if out is not None:
            return out
    out = np.zeros(p.shape[:-1] + (n, n))
    il = np.tril_indices(n)
    out[..., il[0], il[1]] = p
    out[..., il[1], il[0]] = p
    return out


class TaggedArray(np.ndarray):
    """pyscf.lib.tag_array: ndarray carrying ``mo_coeff`` / ``mo_occ`` (or ``ecoul`` / ``exc``)."""


def tag_array(a, **kw):
    t = np.asarray(a).view(TaggedArray)
    for k, v in kw.items():
        setattr(t, k, v)
    return t


# --------------------------------------------------------------------------------------------
# pyscf/df/df_jk.py:get_jk  (with the occupied-orbital K branch and the dense-dm branch)
# --------------------------------------------------------------------------------------------
def df_get_jk(cderi: np.ndarray, dm, with_j=True, with_k=True, blksize: int = 240):
    """``cderi`` is ``[naux, nao(nao+1)/2]`` packed-lower (PySCF ``with_df._cderi`` layout)."""
    dms = np.asarray(dm)
    dm_shape = dms.shape
    nao = dm_shape[-1]
    dms = dms.reshape(-1, nao, nao)
    nset = dms.shape[0]
    naux = cderi.shape[0]
    vj = np.zeros((nset, nao * (nao + 1) // 2))
    vk = np.zeros_like(dms)
    if with_j:
        idx = np.arange(nao)
        dmtril = pack_tril(dms + dms.transpose(0, 2, 1))
        dmtril[:, idx * (idx + 1) // 2 + idx] *= 0.5
    mo_coeff = getattr(dm, "mo_coeff", None)
    orbo = None
    if with_k and mo_coeff is not None:
        mo_coeff = np.asarray(mo_coeff)
        mo_occ = np.asarray(dm.mo_occ)
        nmo = mo_occ.shape[-1]
        mo_coeff = mo_coeff.reshape(-1, nao, nmo)
        mo_occ = mo_occ.reshape(-1, nmo)
        orbo = []
        for k in range(nset):
            sel = mo_occ[k] > 0
            orbo.append(mo_coeff[k][:, sel] * np.sqrt(mo_occ[k][sel]))
    for p0 in range(0, naux, blksize):
        eri1 = cderi[p0 : p0 + blksize]
        if with_j:
            rho = dmtril @ eri1.T
            vj += rho @ eri1
        if with_k:
            b = unpack_tril(eri1, nao)  # [p, n, n]
            for k in range(nset):
                if orbo is not None:
                    if orbo[k].shape[1] > 0:
                        buf1 = (b.reshape(-1, nao) @ orbo[k]).reshape(b.shape[0], nao, -1)  # [p, m, i]
                        buf1 = np.ascontiguousarray(buf1.transpose(1, 0, 2)).reshape(nao, -1)  # [m, (p, i)]
                        vk[k] += buf1 @ buf1.T
                else:
                    buf1 = np.einsum("pmn,nl->pml", b, dms[k], optimize=True)
                    vk[k] += np.einsum("pml,pln->mn", buf1, b, optimize=True)
    vj_out = unpack_tril(vj, nao).reshape(dm_shape) if with_j else None
    vk_out = vk.reshape(dm_shape) if with_k else None
    return vj_out, vk_out


def df_get_jk_occ(cderi: np.ndarray, orbo: list[np.ndarray]):
    """Same contraction from scaled occupied orbitals only (what the CUDA path is fed)."""
    nao = orbo[0].shape[0]
    dms = np.array([c @ c.T for c in orbo])
    mo = np.zeros((len(orbo), nao, max(c.shape[1] for c in orbo)))
    occ = np.zeros((len(orbo), mo.shape[2]))
    for k, c in enumerate(orbo):
        mo[k, :, : c.shape[1]] = c
        occ[k, : c.shape[1]] = 1.0
    return df_get_jk(cderi, tag_array(dms, mo_coeff=mo, mo_occ=occ))


# --------------------------------------------------------------------------------------------
# pyscf/lib/diis.py:DIIS  (in-core; space=6, min_space=1)
# --------------------------------------------------------------------------------------------
class DIIS:
    def __init__(self):
        self.space = 6
        self.min_space = 1
        self._buffer = {}
        self._bookkeep = []
        self._head = 0
        self._H = None
        self._xprev = None
        self._err_vec_touched = False

    def push_err_vec(self, xerr):
        self._err_vec_touched = True
        if self._head >= self.space:
            self._head = 0
        self._buffer["e%d" % self._head] = np.array(xerr).ravel()

    def push_vec(self, x):
        x = np.array(x).ravel()
        while len(self._bookkeep) >= self.space:
            self._bookkeep.pop(0)
        if self._err_vec_touched:
            self._bookkeep.append(self._head)
            self._buffer["x%d" % self._head] = x
            self._head += 1
        elif self._xprev is None:
            self._xprev = x
        else:
            if self._head >= self.space:
                self._head = 0
            self._bookkeep.append(self._head)
            self._buffer["x%d" % self._head] = x
            self._buffer["e%d" % self._head] = x - self._xprev
            self._head += 1

    def get_num_vec(self):
        return len(self._bookkeep)

    def update(self, x, xerr=None):
        if xerr is not None:
            self.push_err_vec(xerr)
        self.push_vec(x)
        nd = self.get_num_vec()
        if nd < self.min_space:
            return x
        dt = self._buffer["e%d" % (self._head - 1)]
        if self._H is None:
            self._H = np.zeros((self.space + 1, self.space + 1))
            self._H[0, 1:] = self._H[1:, 0] = 1
        for i in range(nd):
            tmp = np.dot(dt, self._buffer["e%d" % i])
            self._H[self._head, i + 1] = tmp
            self._H[i + 1, self._head] = tmp
        if self._xprev is None:
            xnew = self.extrapolate(nd)
        else:
            self._xprev = None
            self._xprev = xnew = self.extrapolate(nd)
        return xnew.reshape(np.shape(x))

    def coefficients(self, nd):
        h = self._H[: nd + 1, : nd + 1]
        g = np.zeros(nd + 1)
        g[0] = 1
        w, v = scipy.linalg.eigh(h)
        if np.any(abs(w) < 1e-14):
            idx = abs(w) > 1e-14
            c = np.dot(v[:, idx] * (1.0 / w[idx]), np.dot(v[:, idx].T, g))
        else:
            c = np.linalg.solve(h, g)
        return c

    def extrapolate(self, nd=None):
        if nd is None:
            nd = self.get_num_vec()
        if nd == 0:
            raise RuntimeError("No vector found in DIIS object.")
        c = self.coefficients(nd)
        xnew = None
        for i, ci in enumerate(c[1:]):
            xi = self._buffer["x%d" % i]
            if xnew is None:
                xnew = np.zeros(xi.size)
            xnew += xi * ci
        return xnew


# --------------------------------------------------------------------------------------------
# pyscf/scf/diis.py:CDIIS  (space=8; error vector = Corth^T (S D F - F D S) Corth per spin)
# --------------------------------------------------------------------------------------------
class CDIIS(DIIS):
    def __init__(self, Corth=None):
        super().__init__()
        self.space = 8
        self.Corth = Corth

    def update(self, s, d, f):  # noqa: D102
        errvec = cdiis_err_vec(s, d, f, self.Corth)
        return DIIS.update(self, f, xerr=errvec)


def cdiis_err_vec(s, d, f, Corth=None):
    f = np.asarray(f)
    d = np.asarray(d)
    if f.ndim == 2:
        sdf = s @ d @ f
        if Corth is not None:
            sdf = Corth.T @ sdf @ Corth
        return (sdf.T - sdf).ravel()
    out = []
    for i in range(f.shape[0]):
        sdf = s @ d[i] @ f[i]
        if Corth is not None:
            c = Corth[i] if np.ndim(Corth) == 3 else Corth
            sdf = c.T @ sdf @ c
        out.append((sdf.T - sdf).ravel())
    return np.hstack(out)


# --------------------------------------------------------------------------------------------
# pyscf/scf/uhf.py + hf.py: get_occ, make_rdm1, eig, get_grad
# --------------------------------------------------------------------------------------------
def get_occ_uhf(mo_energy, nelec):
    mo_energy = np.asarray(mo_energy)
    mo_occ = np.zeros_like(mo_energy)
    for s in range(2):
        idx = np.argsort(mo_energy[s].round(9), kind="stable")
        mo_occ[s, idx[: nelec[s]]] = 1
    return mo_occ


def get_occ_rhf(mo_energy, nelectron):
    mo_energy = np.asarray(mo_energy)
    idx = np.argsort(mo_energy.round(9), kind="stable")
    mo_occ = np.zeros_like(mo_energy)
    mo_occ[idx[: nelectron // 2]] = 2
    return mo_occ


def make_rdm1(mo_coeff, mo_occ):
    mo_coeff = np.asarray(mo_coeff)
    mo_occ = np.asarray(mo_occ)
    if mo_coeff.ndim == 2:
        mocc = mo_coeff[:, mo_occ > 0]
        dm = (mocc * mo_occ[mo_occ > 0]) @ mocc.T
    else:
        dm = []
        for s in range(mo_coeff.shape[0]):
            mocc = mo_coeff[s][:, mo_occ[s] > 0]
            dm.append((mocc * mo_occ[s][mo_occ[s] > 0]) @ mocc.T)
        dm = np.array(dm)
    return tag_array(dm, mo_coeff=mo_coeff, mo_occ=mo_occ)


def eig_generalized(f, s):
    """pyscf.scf.hf.eig: scipy.linalg.eigh(f, s) + the largest-|c| component made positive."""
    e, c = scipy.linalg.eigh(f, s)
    idx = np.argmax(abs(c), axis=0)
    c[:, c[idx, np.arange(len(e))] < 0] *= -1
    return e, c


def get_grad_uhf(mo_coeff, mo_occ, fock):
    g = []
    for s in range(2):
        occ = mo_occ[s] > 0
        g.append((mo_coeff[s][:, ~occ].T @ (fock[s] @ mo_coeff[s][:, occ])).ravel())
    return np.hstack(g)


def get_grad_rhf(mo_coeff, mo_occ, fock):
    occ = mo_occ > 0
    return (mo_coeff[:, ~occ].T @ (fock @ mo_coeff[:, occ])).ravel() * 2


# --------------------------------------------------------------------------------------------
# A duck-typed DF SCF object satisfying the protocol of SURVEY.md §8(b)
# --------------------------------------------------------------------------------------------
class _Mol:
    def __init__(self, nelec, e_nuc=0.0):
        self.nelec = tuple(nelec)
        self.nelectron = int(sum(nelec))
        self.spin = nelec[0] - nelec[1]
        self._e_nuc = e_nuc

    def energy_nuc(self):
        return self._e_nuc


class DFSCFBase:
    """DF-HF object over explicit tensors (S, hcore, packed cderi): the oracle's ``mf.density_fit()``."""

    unrestricted = True

    def __init__(self, ovlp, hcore, cderi, nelec, e_nuc=0.0, max_cycle=50, conv_tol=1e-9):
        self._s = np.asarray(ovlp)
        self._h = np.asarray(hcore)
        # a row-streamed tensor (oracle/streamed.py) stands in for pyscf's on-disk cderi at sizes that do not fit RAM
        self.cderi = cderi if getattr(cderi, "rows_are_streamed", False) else np.asarray(cderi)
        self.mol = _Mol(nelec, e_nuc)
        self.max_cycle = max_cycle
        self.conv_tol = conv_tol
        self.mo_coeff = self.mo_occ = self.mo_energy = None
        self.e_tot = None
        self.converged = False
        self.scf_summary = {}
        self.n_jk_builds = 0

    @property
    def nelec(self):
        return self.mol.nelec

    def get_ovlp(self, *a):
        return self._s

    def get_hcore(self, *a):
        return self._h

    def energy_nuc(self):
        return self.mol.energy_nuc()

    def get_jk(self, mol=None, dm=None, hermi=1, with_j=True, with_k=True):
        self.n_jk_builds += 1
        return df_get_jk(self.cderi, dm, with_j, with_k)

    def get_j(self, mol=None, dm=None, hermi=1):
        return self.get_jk(mol, dm, hermi, with_k=False)[0]

    def make_rdm1(self, mo_coeff=None, mo_occ=None):
        if mo_coeff is None:
            mo_coeff = self.mo_coeff
        if mo_occ is None:
            mo_occ = self.mo_occ
        return make_rdm1(mo_coeff, mo_occ)

    def eig(self, f, s):
        f = np.asarray(f)
        if f.ndim == 2:
            return eig_generalized(f, s)
        ea, ca = eig_generalized(f[0], s)
        eb, cb = eig_generalized(f[1], s)
        return np.array((ea, eb)), np.array((ca, cb))

    def energy_tot(self, dm=None, h1e=None, vhf=None):
        return self.energy_elec(dm, h1e, vhf)[0] + self.energy_nuc()


class DFUHF(DFSCFBase):
    unrestricted = True

    def get_veff(self, mol=None, dm=None, dm_last=0, vhf_last=0, hermi=1):
        if dm is None:
            dm = self.make_rdm1()
        if isinstance(dm, np.ndarray) and dm.ndim == 2:
            dm = np.asarray((dm * 0.5, dm * 0.5))
        vj, vk = self.get_jk(mol, dm, hermi)
        return vj[0] + vj[1] - vk

    def get_occ(self, mo_energy=None, mo_coeff=None):
        return get_occ_uhf(mo_energy, self.nelec)

    def get_fock(self, h1e=None, s1e=None, vhf=None, dm=None, cycle=-1, diis=None, diis_start_cycle=1):
        if h1e is None:
            h1e = self.get_hcore()
        if vhf is None:
            vhf = self.get_veff(dm=dm)
        f = np.asarray(h1e) + vhf
        if f.ndim == 2:
            f = np.array((f, f))
        if cycle < 0 and diis is None:
            return f
        if diis is not None and cycle >= diis_start_cycle:
            f = diis.update(s1e, dm, f)
        return f

    def get_grad(self, mo_coeff, mo_occ, fock):
        return get_grad_uhf(mo_coeff, mo_occ, fock)

    def energy_elec(self, dm=None, h1e=None, vhf=None):
        """pyscf.scf.uhf.energy_elec for 2-D hcore (the 3-D variant is the reference's own function)."""
        if dm is None:
            dm = self.make_rdm1()
        if h1e is None:
            h1e = self.get_hcore()
        if vhf is None:
            vhf = self.get_veff(dm=dm)
        h1e = np.asarray(h1e)
        if h1e.ndim == 2:
            h1e = (h1e, h1e)
        e1 = np.einsum("ij,ji->", h1e[0], dm[0]) + np.einsum("ij,ji->", h1e[1], dm[1])
        e_coul = 0.5 * (np.einsum("ij,ji->", vhf[0], dm[0]) + np.einsum("ij,ji->", vhf[1], dm[1]))
        return e1 + e_coul, e_coul


class DFRHF(DFSCFBase):
    unrestricted = False

    def get_veff(self, mol=None, dm=None, dm_last=0, vhf_last=0, hermi=1):
        if dm is None:
            dm = self.make_rdm1()
        vj, vk = self.get_jk(mol, dm, hermi)
        return vj - vk * 0.5

    def get_occ(self, mo_energy=None, mo_coeff=None):
        return get_occ_rhf(mo_energy, self.mol.nelectron)

    def get_fock(self, h1e=None, s1e=None, vhf=None, dm=None, cycle=-1, diis=None, diis_start_cycle=1):
        if h1e is None:
            h1e = self.get_hcore()
        if vhf is None:
            vhf = self.get_veff(dm=dm)
        f = h1e + vhf
        if cycle < 0 and diis is None:
            return f
        if diis is not None and cycle >= diis_start_cycle:
            f = diis.update(s1e, dm, f)
        return f

    def get_grad(self, mo_coeff, mo_occ, fock):
        return get_grad_rhf(mo_coeff, mo_occ, fock)

    def energy_elec(self, dm=None, h1e=None, vhf=None):
        if dm is None:
            dm = self.make_rdm1()
        if h1e is None:
            h1e = self.get_hcore()
        if vhf is None:
            vhf = self.get_veff(dm=dm)
        e1 = np.einsum("ij,ji->", h1e, dm)
        e_coul = 0.5 * np.einsum("ij,ji->", vhf, dm)
        return e1 + e_coul, e_coul


# --------------------------------------------------------------------------------------------
# pyscf/scf/hf.py:kernel  (control flow of the mu-shift path, nbed/driver.py:533)
# --------------------------------------------------------------------------------------------
def scf_kernel(mf, conv_tol=1e-10, conv_tol_grad=None, dm0=None, conv_check=True, use_diis=True, trace=None):
    """Returns (scf_conv, e_tot, mo_energy, mo_coeff, mo_occ) and writes them onto ``mf`` like ``mf.kernel()``.

    ``dm0`` must be supplied: PySCF's default 'minao' guess needs atomic basis data that is not available.
    ``trace`` (list) receives per-cycle (e_tot, norm_gorb, norm_ddm) for iterate-level parity tests.
    """
    if conv_tol_grad is None:
        conv_tol_grad = np.sqrt(conv_tol)
    s1e = mf.get_ovlp()
    dm = dm0
    h1e = mf.get_hcore()
    vhf = mf.get_veff(dm=dm)
    e_tot = mf.energy_tot(dm, h1e, vhf)
    scf_conv = False
    mo_energy = mo_coeff = mo_occ = None
    mf_diis = None
    if use_diis:
        mf_diis = CDIIS()
        fock = mf.get_fock(h1e, s1e, vhf, dm)
        _, mf_diis.Corth = mf.eig(fock, s1e)
    cycle = -1
    for cycle in range(mf.max_cycle):
        dm_last = dm
        last_hf_e = e_tot
        fock = mf.get_fock(h1e, s1e, vhf, dm, cycle, mf_diis)
        mo_energy, mo_coeff = mf.eig(fock, s1e)
        mo_occ = mf.get_occ(mo_energy, mo_coeff)
        dm = mf.make_rdm1(mo_coeff, mo_occ)
        vhf = mf.get_veff(dm=dm)
        e_tot = mf.energy_tot(dm, h1e, vhf)
        fock = mf.get_fock(h1e, s1e, vhf, dm)
        norm_gorb = np.linalg.norm(mf.get_grad(mo_coeff, mo_occ, fock))
        norm_ddm = np.linalg.norm(np.asarray(dm) - np.asarray(dm_last))
        if trace is not None:
            trace.append((float(e_tot), float(norm_gorb), float(norm_ddm)))
        if abs(e_tot - last_hf_e) < conv_tol and norm_gorb < conv_tol_grad:
            scf_conv = True
        if scf_conv:
            break
    mf.cycles = cycle + 1
    if scf_conv and conv_check:
        mo_energy, mo_coeff = mf.eig(fock, s1e)
        mo_occ = mf.get_occ(mo_energy, mo_coeff)
        dm, dm_last = mf.make_rdm1(mo_coeff, mo_occ), dm
        vhf = mf.get_veff(dm=dm)
        e_tot, last_hf_e = mf.energy_tot(dm, h1e, vhf), e_tot
        fock = mf.get_fock(h1e, s1e, vhf, dm)
        norm_gorb = np.linalg.norm(mf.get_grad(mo_coeff, mo_occ, fock))
        norm_ddm = np.linalg.norm(np.asarray(dm) - np.asarray(dm_last))
        if trace is not None:
            trace.append((float(e_tot), float(norm_gorb), float(norm_ddm)))
        scf_conv = bool(abs(e_tot - last_hf_e) < conv_tol * 10 or norm_gorb < conv_tol_grad * 3)
    mf.converged, mf.e_tot = scf_conv, e_tot
    mf.mo_energy, mf.mo_coeff, mf.mo_occ = mo_energy, mo_coeff, mo_occ
    return scf_conv, e_tot, mo_energy, mo_coeff, mo_occ


# --------------------------------------------------------------------------------------------
# pyscf/df/df_ao2mo.py + pyscf/ao2mo: kernel / restore
# --------------------------------------------------------------------------------------------
def df_half_transformed(cderi, ci, cj):
    """L[P, i, j] = sum_mn ci[m,i] B[P,m,n] cj[n,j]."""
    nao = ci.shape[0]
    out = np.empty((cderi.shape[0], ci.shape[1], cj.shape[1]))
    for p0 in range(0, cderi.shape[0], 128):
        b = unpack_tril(cderi[p0 : p0 + 128], nao)
        out[p0 : p0 + 128] = np.einsum("mi,pmn,nj->pij", ci, b, cj, optimize=True)
    return out


def ao2mo_kernel(cderi, mo_coeffs):
    """``ao2mo.kernel(mol, (C1,C2,C3,C4))`` on the DF tensor -> (ij|kl) as a 2-D matrix.

    PySCF packs (s4) when C1 is C2 and C3 is C4; ``ao2mo_restore(1, ...)`` undoes it, so the dense
    (n1*n2, n3*n4) matrix carries the same information and is what is returned here, tagged with shapes.
    """
    if isinstance(mo_coeffs, np.ndarray) and mo_coeffs.ndim == 2:
        mo_coeffs = (mo_coeffs,) * 4
    c1, c2, c3, c4 = mo_coeffs
    lij = df_half_transformed(cderi, c1, c2)
    lkl = lij if (c3 is c1 and c4 is c2) else df_half_transformed(cderi, c3, c4)
    naux = cderi.shape[0]
    eri = lij.reshape(naux, -1).T @ lkl.reshape(naux, -1)
    compact = (c1 is c2) and (c3 is c4)
    if compact:
        n12, n34 = c1.shape[1], c3.shape[1]
        i1 = np.tril_indices(n12)
        i3 = np.tril_indices(n34)
        eri4 = eri.reshape(n12, n12, n34, n34)
        return eri4[i1[0], i1[1]][:, i3[0], i3[1]]
    return eri


def ao2mo_restore(symm, eri, norb):
    """``ao2mo.restore(1, eri, norb)``: s4-packed or dense 2-D -> (norb,)*4."""
    assert symm == 1
    npair = norb * (norb + 1) // 2
    eri = np.asarray(eri)
    if eri.shape == (npair, npair):
        return unpack_tril(unpack_tril(eri, norb).transpose(1, 2, 0), norb).transpose(2, 3, 0, 1).copy()
    return eri.reshape(norb, norb, norb, norb)


# --------------------------------------------------------------------------------------------
# pyscf/df/incore.py:cholesky_eri, and the exact-ERI shortcut for tiny systems
# --------------------------------------------------------------------------------------------
def cholesky_eri_exact(eri4: np.ndarray, tol: float = 1e-14) -> np.ndarray:
    """Full-rank factorisation of an exact 4-index ERI into packed 'cderi' rows.

    (mn|ls) = sum_P B[P,mn] B[P,ls] to ~1e-14, so the DF kernels reproduce exact-ERI results
    (the way to reach 'same as the unmodified reference' for the tiny configs).
    """
    n = eri4.shape[0]
    il = np.tril_indices(n)
    m = eri4[il[0], il[1]][:, il[0], il[1]]  # (npair, npair), symmetric PSD
    w, v = np.linalg.eigh(m)
    keep = w > tol * w.max()
    return np.ascontiguousarray((v[:, keep] * np.sqrt(w[keep])).T)
