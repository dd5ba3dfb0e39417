"""ctypes binding of oracle/c_kernels.c (plain-C restatement of the DF J/K contraction).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib = None


def load(build: bool = True):
    """Returns the library or None when it cannot be built (callers fall back to the NumPy restatement)."""
    global _lib
    if _lib is not None:
        return _lib
    stale = os.path.exists(_LIB) and os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "c_kernels.c"))
    if (stale or not os.path.exists(_LIB)) and build:
        try:
            subprocess.run(["make", "-s", "-C", _HERE], check=True, capture_output=True, timeout=120)
        except Exception:
            return None
    if not os.path.exists(_LIB):
        return None
    lib = C.CDLL(_LIB)
    lib.oracle_unpack_tril.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int]
    lib.oracle_unpack_tril.restype = None
    lib.oracle_df_jk_occ.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_df_jk_occ.restype = None
    if hasattr(lib, "oracle_synth_rows"):
        lib.oracle_synth_rows.argtypes = [C.c_ulonglong, C.c_int, C.c_double, C.c_long, C.c_long, C.c_void_p]
        lib.oracle_synth_rows.restype = None
    _lib = lib
    return lib


def unpack_tril(packed: np.ndarray, n: int) -> np.ndarray | None:
    lib = load()
    if lib is None:
        return None
    packed = np.ascontiguousarray(packed, dtype=np.float64)
    rows = packed.reshape(-1, packed.shape[-1])
    out = np.empty((rows.shape[0], n, n))
    lib.oracle_unpack_tril(rows.ctypes.data, out.ctypes.data, rows.shape[0], n)
    return out.reshape(packed.shape[:-1] + (n, n))


def df_jk_occ(cderi: np.ndarray, orbs):
    """(vj, vk) of shape (nset, n, n) from scaled occupied orbital blocks (n, ncol)."""
    lib = load()
    if lib is None:
        raise RuntimeError("oracle C library is not built (make -C oracle)")
    cderi = np.ascontiguousarray(cderi, dtype=np.float64)
    n = orbs[0].shape[0]
    ncol = np.array([o.shape[1] for o in orbs], dtype=np.int32)
    flat = np.concatenate([np.ascontiguousarray(o, dtype=np.float64).ravel() for o in orbs]) if ncol.sum() else np.zeros(1)
    vj = np.empty((len(orbs), n, n))
    vk = np.empty((len(orbs), n, n))
    lib.oracle_df_jk_occ(cderi.ctypes.data, cderi.shape[0], n, len(orbs), ncol.ctypes.data, flat.ctypes.data,
                         vj.ctypes.data, vk.ctypes.data)
    return vj, vk


def synth_rows(seed: int, n: int, scale: float, row0: int, nrows: int) -> np.ndarray | None:
    """Rows [row0, row0 + nrows) of the synthetic packed tensor (bit-identical to synthetic.synth_cderi_rows)."""
    lib = load()
    if lib is None or not hasattr(lib, "oracle_synth_rows"):
        return None
    out = np.empty((nrows, n * (n + 1) // 2))
    lib.oracle_synth_rows(int(seed), int(n), float(scale), int(row0), int(nrows), out.ctypes.data)
    return out
