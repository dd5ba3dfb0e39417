"""Host-side mirror of ``nbed.ham_builder.HamiltonianBuilder`` (/root/reference nbed/ham_builder.py:17-285) with
the integral transforms on the GPU (``nbd_ao2mo``, ``nbd_one_body``, ``nbd_spinorb_from_spatial``)."""
from __future__ import annotations

import numpy as np

from .scf import B200RHF, _B200SCF

EQ_TOLERANCE = 1e-8  # openfermion.config.EQ_TOLERANCE (openfermion 1.7.1), nbed/ham_builder.py:8,213-214


class HamiltonianBuilderError(Exception):
    """nbed/exceptions.py"""


class HamiltonianBuilder:
    def __init__(self, scf_method, constant_e_shift: float = 0, n_frozen_core: int = 0, n_frozen_virt: int = 0):
        if not isinstance(scf_method, _B200SCF):
            raise TypeError("HamiltonianBuilder needs a B200RHF / B200UHF SCF object (no CPU fallback)")
        self.scf_method = scf_method
        self.constant_e_shift = constant_e_shift
        self.n_frozen_core = n_frozen_core
        self.n_frozen_virt = n_frozen_virt
        self._restricted = isinstance(scf_method, B200RHF)  # :39
        occ = np.asarray(scf_method.mo_occ)
        if occ.ndim == 1:
            self.occupancy = occ
        elif occ.ndim == 2:
            self.occupancy = np.vstack((occ[0], occ[1]))
        else:
            raise HamiltonianBuilderError("occupancy dimension error")  # :49

    @property
    def _one_body_integrals(self):  # :53-96
        c = np.asarray(self.scf_method.mo_coeff)
        hcore = np.asarray(self.scf_method.get_hcore())
        ctx = self.scf_method.ctx
        if not self._restricted:
            return ctx.one_body(hcore, c[0], c[1])
        return ctx.one_body(hcore, c)

    @property
    def _two_body_integrals(self):  # :98-156
        c = np.asarray(self.scf_method.mo_coeff)
        ctx = self.scf_method.ctx
        if not self._restricted:
            if c[0].shape[1] != c[1].shape[1]:
                raise HamiltonianBuilderError("Must localize the same number of alpha and beta orbitals.")  # :109-112
            return ctx.ao2mo(c[0], c[1])
        return ctx.ao2mo(c)

    def _spinorb_from_spatial(self, one_body_integrals, two_body_integrals, two_body_scale: float = 1.0):  # :158-216
        return self.scf_method.ctx.spinorb_from_spatial(one_body_integrals, two_body_integrals, EQ_TOLERANCE,
                                                        two_body_scale)

    def build(self):  # :218-254
        if self.n_frozen_virt != 0:
            self.scf_method = reduce_virtuals(self.scf_method, self.n_frozen_virt)
        # one fused device call: one-body + four two-body blocks + spin-orbital scatter (EQ_TOLERANCE, the 0.5 of :254)
        c = np.asarray(self.scf_method.mo_coeff)
        hcore = np.asarray(self.scf_method.get_hcore())
        ctx = self.scf_method.ctx
        if not self._restricted:
            if c[0].shape[1] != c[1].shape[1]:
                raise HamiltonianBuilderError("Must localize the same number of alpha and beta orbitals.")  # :109-112
            h1, h2 = ctx.build_hamiltonian(hcore, c[0], c[1], EQ_TOLERANCE, 0.5)
        else:
            h1, h2 = ctx.build_hamiltonian(hcore, c, None, EQ_TOLERANCE, 0.5)
        return self.constant_e_shift, h1, h2


def reduce_virtuals(scf_method, n_frozen_virt: int):  # :257-285
    reduced = scf_method.copy()
    if n_frozen_virt <= 0:
        return reduced
    if n_frozen_virt >= np.count_nonzero(reduced.mo_occ):
        raise ValueError("Atempting to reduce virtual space by more than exist.")
    if reduced.unrestricted:
        reduced.mo_coeff = np.asarray(reduced.mo_coeff)[:, :, :-n_frozen_virt]
        reduced.mo_occ = np.asarray(reduced.mo_occ)[:, :-n_frozen_virt]
    else:
        reduced.mo_coeff = np.asarray(reduced.mo_coeff)[:, :-n_frozen_virt]
        reduced.mo_occ = np.asarray(reduced.mo_occ)[:-n_frozen_virt]
    return reduced
