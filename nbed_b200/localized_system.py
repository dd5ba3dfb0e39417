"""Mirror of ``nbed.localizers.system.LocalizedSystem`` (/root/reference nbed/localizers/system.py:8-36): the data carrier
the localizers hand to the embedding drivers (``c_active`` / ``c_enviro`` / ``c_loc_occ`` and the densities
``C C^T`` derived from them in ``__post_init__``).

Same fields, same derived attributes.  One addition that the GPU path exploits: the derived densities are
``TaggedArray``s carrying their own factor (``dm.factor``, the orbital block they were built from), so that
``huzinaga_scf`` can hand ``c_enviro`` to ``nbd_scf_set_env_orbitals`` and run the projector product
``F gamma S`` as two rank-r GEMMs instead of a dense n^3 one.  A plain ndarray works everywhere a tagged one does.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from numpy.typing import NDArray


class FactoredDensity(np.ndarray):
    """ndarray (.., n, n) with ``factor`` (.., n, r) such that ``self == factor @ factor.swapaxes(-1, -2)``."""


def _density(c: NDArray) -> NDArray:
    c = np.asarray(c, dtype=np.float64)
    dm = (c @ c.swapaxes(-1, -2)).view(FactoredDensity)
    dm.factor = c
    return dm


@dataclass
class LocalizedSystem:
    """Required data from localized system (field names and meaning as in the reference)."""

    active_mo_inds: NDArray
    enviro_mo_inds: NDArray
    c_active: NDArray
    c_enviro: NDArray
    c_loc_occ: NDArray
    c_loc_virt: NDArray | None = None
    dm_active: NDArray = field(init=False)
    dm_enviro: NDArray = field(init=False)
    dm_loc_occ: NDArray = field(init=False)

    def __post_init__(self):
        """Post init for derived attributes (system.py:32-36)."""
        self.dm_active = _density(self.c_active)
        self.dm_enviro = _density(self.c_enviro)
        self.dm_loc_occ = _density(self.c_loc_occ)
