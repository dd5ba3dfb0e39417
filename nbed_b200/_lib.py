"""ctypes binding of the C-ABI library (include/nbed_b200.h).  No CPU fallback: if the CUDA library is
missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnbed_b200.so")

NBD_HUZINAGA = 0
NBD_MU_SHIFT = 1

# every symbol include/nbed_b200.h declares (tests check that the built library exports all of them)
EXPORTS = [
    "nbd_version", "nbd_create", "nbd_destroy", "nbd_last_error", "nbd_set_option", "nbd_timer_ms",
    "nbd_launch_count", "nbd_host_alloc", "nbd_host_free", "nbd_comm_unique_id", "nbd_comm_init", "nbd_cderi_alloc", "nbd_cderi_upload",
    "nbd_cderi_synth", "nbd_cderi_download", "nbd_basis_dims", "nbd_int3c2e", "nbd_cderi_from_basis", "nbd_xc_setup", "nbd_xc_nr_uks", "nbd_scf_set_xc", "nbd_jk", "nbd_jk_dm", "nbd_scf_setup", "nbd_scf_set_virtual_projector", "nbd_scf_set_env_orbitals", "nbd_huzinaga_scf",
    "nbd_mu_scf", "nbd_scf_bench_init", "nbd_scf_bench_iteration", "nbd_ao2mo", "nbd_one_body",
    "nbd_spinorb_from_spatial", "nbd_build_hamiltonian",
]


class NbdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nbed_b200 error {code}: {msg}")
        self.code = code


class ScfResult(C.Structure):
    _fields_ = [
        ("converged", C.c_int),
        ("cycles", C.c_int),
        ("e_tot", C.c_double),
        ("energy", C.c_double * 2),
        ("norm_ddm", C.c_double),
        ("norm_grad", C.c_double),
    ]


_lib = None


def load() -> C.CDLL:
    """Load (once) the in-tree CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m nbed_b200.build` (nvcc, sm_100a). "
            "nbed_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    P, I, D, L = C.c_void_p, C.c_int, C.c_double, C.c_long
    sig = {
        "nbd_version": (I, []),
        "nbd_create": (I, [C.POINTER(P), I]),
        "nbd_destroy": (I, [P]),
        "nbd_last_error": (C.c_char_p, [P]),
        "nbd_set_option": (I, [P, C.c_char_p, L]),
        "nbd_timer_ms": (D, [P, C.c_char_p]),
        "nbd_launch_count": (L, [P]),
        "nbd_host_alloc": (P, [C.c_size_t]),
        "nbd_host_free": (None, [P]),
        "nbd_comm_unique_id": (I, [P]),
        "nbd_comm_init": (I, [P, P, I, I]),
        "nbd_cderi_alloc": (I, [P, I, I]),
        "nbd_cderi_upload": (I, [P, P, I, I]),
        "nbd_cderi_synth": (I, [P, C.c_ulonglong, D, I]),
        "nbd_cderi_download": (I, [P, P, I, I]),
        "nbd_basis_dims": (I, [P, I, I, P, P]),
        "nbd_xc_setup": (I, [P, I, P, I, P, I, P, I, I, P, P]),
        "nbd_xc_nr_uks": (I, [P, P, P, P, P]),
        "nbd_scf_set_xc": (I, [P, I]),
        "nbd_int3c2e": (I, [P, P, I, P, I, P, I, I, P, P]),
        "nbd_cderi_from_basis": (I, [P, P, I, P, I, P, I, I, I, I]),
        "nbd_jk": (I, [P, I, P, P, P, P, P]),
        "nbd_jk_dm": (I, [P, I, P, P, P]),
        "nbd_scf_setup": (I, [P, I, P, P, P, P, P, I, D]),
        "nbd_scf_set_virtual_projector": (I, [P, P]),
        "nbd_scf_set_env_orbitals": (I, [P, I, P]),
        "nbd_huzinaga_scf": (I, [P, I, D, D, I, P, P, P, P, P, P, C.POINTER(ScfResult)]),
        "nbd_mu_scf": (I, [P, I, D, D, P, P, P, P, P, P, P, C.POINTER(ScfResult)]),
        "nbd_scf_bench_init": (I, [P]),
        "nbd_scf_bench_iteration": (I, [P, I, P, P]),
        "nbd_ao2mo": (I, [P, I, P, P, P]),
        "nbd_one_body": (I, [P, I, I, P, P, P, P]),
        "nbd_spinorb_from_spatial": (I, [P, I, P, P, D, D, P, P]),
        "nbd_build_hamiltonian": (I, [P, I, I, P, P, P, D, D, P, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Pointer to a C-contiguous float64 / int32 array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)


# Page-locked result buffers.  A device -> host copy into a fresh ``np.empty`` array pays for first-touch page faults
# and for the staging pipeline of pageable memory (~9 ms per 82 MB ao2mo result, ~10 ms per Huzinaga SCF call at
# n = 1376); page-locking a new buffer per call costs even more.  So result arrays of 1 MiB and more are carved from
# blocks that return to a pool when the array (and every view of it) has been garbage collected.
_POOL: dict = {}
_POOL_STATE = {"bytes": 0, "cap": 2 << 30}
_POOL_MIN_BYTES = 1 << 20


def _pool_release(p: int, nbytes: int) -> None:
    if _POOL_STATE["bytes"] + nbytes <= _POOL_STATE["cap"]:
        _POOL.setdefault(nbytes, []).append(p)
        _POOL_STATE["bytes"] += nbytes
    else:
        load().nbd_host_free(p)


def pinned_empty(shape) -> np.ndarray:
    """float64 array in page-locked host memory, recycled through a pool once the array is garbage collected.
    Falls back to a pageable array only if the allocation itself fails (results are identical, the copy is slower)."""
    import weakref

    lib = load()
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    n = int(np.prod(shape)) if shape else 1
    nbytes = n * 8
    free = _POOL.get(nbytes)
    if free:
        p = free.pop()
        _POOL_STATE["bytes"] -= nbytes
    else:
        p = lib.nbd_host_alloc(nbytes)
    if not p:
        return np.empty(shape)
    buf = (C.c_double * n).from_address(p)
    arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
    weakref.finalize(buf, _pool_release, p, nbytes)
    return arr


def result_empty(shape) -> np.ndarray:
    """Output array of a device call: page-locked (pooled) from 1 MiB, plain ``np.empty`` below."""
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    if int(np.prod(shape)) * 8 < _POOL_MIN_BYTES:
        return np.empty(shape)
    return pinned_empty(shape)


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)
