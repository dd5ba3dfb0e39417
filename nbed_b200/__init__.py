"""nbed_b200 — B200-native (sm_100a) implementation of Nbed's embedded-SCF Fock build and active-space AO->MO
transform behind the reference's own operator interface.  See DESIGN.md / INTEGRATION.md."""
from .backend import NBD_HUZINAGA, NBD_MU_SHIFT, B200Context, NbdError  # noqa: F401
from .ham_builder import HamiltonianBuilder, reduce_virtuals  # noqa: F401
from .scf import (B200RHF, B200RKS, B200UHF, B200UKS, energy_elec, get_huzinaga_operator, huzinaga_embed, huzinaga_scf,  # noqa: F401
                  mu_embed)
from .localized_system import LocalizedSystem  # noqa: F401,E402
