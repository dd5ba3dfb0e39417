"""Stub ``pyscf`` / ``openfermion`` modules so the reference's pure-NumPy hot-path code imports UNMODIFIED.

TEST INFRASTRUCTURE, dev container only (``/root/reference`` does not exist on the GPU box).  The stub
classes are backed by ``oracle.pyscf_restatement`` so that, e.g., ``nbed.scf.huzinaga_scf.huzinaga_scf``
(reference nbed/scf/huzinaga_scf.py:93-206) runs its own loop code against the NumPy DF-J/K.
Import surface follows SURVEY.md Appendix C.
"""
from __future__ import annotations

import os
import sys
import types

from . import pyscf_restatement as ps
from . import xc_restatement as xcr

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nbed"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class StreamObject:
    pass


class _RHF(ps.DFRHF, StreamObject):
    pass


class _UHF(ps.DFUHF, StreamObject):
    pass


class _RKS(xcr.DFRKS, StreamObject):
    """Stands in for pyscf.dft.rks.RKS (the object type of the reference's tests/test_scf.py:19-40)."""


class _UKS(xcr.DFUKS, StreamObject):
    """Stands in for pyscf.dft.uks.UKS: passes the reference's isinstance test (huzinaga_scf.py:176) and provides
    get_veff with the .ecoul / .exc tags that calculate_ks_energy reads (:55-56)."""


class _ROHF(StreamObject):
    pass


class _ROKS(StreamObject):
    pass


class _KohnShamDFT:
    pass


class _Mole:
    pass


def install():
    """Insert the stub modules and put the reference on sys.path.  Returns the imported ``nbed``."""
    if not reference_available():
        raise RuntimeError("reference tree not present; stubs are only usable in the dev container")
    if "pyscf" in sys.modules and not getattr(sys.modules["pyscf"], "_nbed_b200_stub", False):
        raise RuntimeError("a real pyscf is importable: use it instead of the stubs")
    lib_diis = _mod("pyscf.lib.diis", DIIS=ps.DIIS)
    lib_misc = _mod("pyscf.lib.misc", StreamObject=StreamObject)
    lib = _mod("pyscf.lib", StreamObject=StreamObject, diis=lib_diis, misc=lib_misc, tag_array=ps.tag_array)
    scf_rhf = _mod("pyscf.scf.rhf", RHF=_RHF)
    scf_hf = _mod("pyscf.scf.hf", RHF=_RHF)
    scf_uhf = _mod("pyscf.scf.uhf", UHF=_UHF)
    scf = _mod("pyscf.scf", rhf=scf_rhf, hf=scf_hf, uhf=scf_uhf, RHF=_RHF, UHF=_UHF, ROHF=_ROHF)
    dft_rks = _mod("pyscf.dft.rks", RKS=_RKS)
    dft_uks = _mod("pyscf.dft.uks", UKS=_UKS)
    dft = _mod("pyscf.dft", rks=dft_rks, uks=dft_uks, RKS=_RKS, UKS=_UKS, ROKS=_ROKS, KohnShamDFT=_KohnShamDFT)
    gto = _mod("pyscf.gto", Mole=_Mole, mole=_Mole, intor_cross=None)
    ao2mo = _mod("pyscf.ao2mo", kernel=None, restore=ps.ao2mo_restore)
    cc = _mod("pyscf.cc", CCSD=type("CCSD", (), {}))
    fci = _mod("pyscf.fci", FCI=type("FCI", (), {}))
    qmmm = _mod("pyscf.qmmm", mm_charge=None)
    lo_vvo = _mod("pyscf.lo.vvo", vvo=None)
    lo_boys = _mod("pyscf.lo.boys", Boys=None)
    lo_iao = _mod("pyscf.lo.iao", iao=None)
    lo_ibo = _mod("pyscf.lo.ibo", ibo=None)
    lo = _mod("pyscf.lo", PipekMezey=None, boys=lo_boys, iao=lo_iao, ibo=lo_ibo, vvo=lo_vvo, vec_lowdin=None)
    _mod("pyscf", lib=lib, scf=scf, dft=dft, ao2mo=ao2mo, gto=gto, cc=cc, fci=fci, qmmm=qmmm, lo=lo,
         _nbed_b200_stub=True)
    of_config = _mod("openfermion.config", EQ_TOLERANCE=1e-8)
    of_pubchem = _mod("openfermion.chem.pubchem", geometry_from_pubchem=None)
    of_chem = _mod("openfermion.chem", pubchem=of_pubchem)
    _mod("openfermion", config=of_config, chem=of_chem)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    try:
        # ``import nbed`` truncates ./.nbed.log (reference nbed/__init__.py:9): keep that out of the repo.
        import tempfile

        os.chdir(tempfile.gettempdir())
        import nbed  # noqa: F401
    finally:
        os.chdir(cwd)
    return sys.modules["nbed"]


def make_scf(kind: str, ovlp, hcore, cderi, nelec, **kw):
    """A stub-typed SCF object that passes the reference's ``isinstance`` checks (huzinaga_scf.py:176,181)."""
    cls = {"rhf": _RHF, "uhf": _UHF, "uks": _UKS, "rks": _RKS}[kind]
    return cls(ovlp, hcore, cderi, nelec, **kw)
