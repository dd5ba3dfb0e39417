"""Builds the C-ABI CUDA library in-tree (``nbed_b200/libnbed_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU.  The library links cuSOLVER (the eigensolvers the north star names)
and resolves NCCL at run time with dlopen, so it loads on hosts without NCCL.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnbed_b200.so")
SOURCES = ["nbed_b200.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))  # every header the one TU includes


def _nccl_include() -> str:
    cands = [os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "nccl", "include"), "/usr/include",
             "/usr/local/cuda/include"]
    for c in cands:
        if os.path.exists(os.path.join(c, "nccl.h")):
            return c
    raise RuntimeError("nccl.h not found (needed for type declarations only)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "nbed_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    cmd = [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-I", _nccl_include(), "-I", os.path.join(HERE, "..", "include"),
        "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcusolver", "-lcublas", "-ldl", "-Xlinker", "-rpath=" + cuda_lib]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    tmp = LIB + ".tmp"
    cmd[cmd.index("-o") + 1] = tmp
    subprocess.run(cmd, check=True)  # raises on any compile error; the old library is only replaced on success
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
