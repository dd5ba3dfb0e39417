"""Row-streamed 3-centre tensor for the oracle.  TEST INFRASTRUCTURE.

At BASELINE config 4 ((H2O)32 / def2-TZVP shape: nao = 1376, naux = 4128) the packed tensor is 31 GB, so the oracle
cannot hold it as one array the way the small tests do.  PySCF itself streams it: ``pyscf/df/df_jk.py:get_jk`` loops
``for eri1 in dfobj.loop(blksize)`` over aux-row blocks read from disk.  ``SyntheticCderi`` plays the part of that
on-disk tensor: slicing it by rows regenerates the block from the counter-based hash of nbed_b200/synthetic.py
(bit-identical on host and device), so ``pyscf_restatement.df_get_jk`` / ``df_half_transformed`` - which only ever
take ``cderi[p0:p1]`` and ``cderi.shape[0]`` - run unchanged at the full size.
"""
from __future__ import annotations

import time

import numpy as np

from nbed_b200 import synthetic as syn

from . import c_binding


class SyntheticCderi:
    """Duck-types the two things the oracle uses of the packed ``[naux, nao(nao+1)/2]`` array."""

    rows_are_streamed = True

    def __init__(self, seed: int, n: int, naux: int, scale: float):
        self.seed, self.n, self.naux, self.scale = int(seed), int(n), int(naux), float(scale)
        self.shape = (self.naux, self.n * (self.n + 1) // 2)
        self.rows_generated = 0
        self.gen_seconds = 0.0  # time spent regenerating rows (bench.py subtracts it from the CPU arm)

    def __len__(self):
        return self.naux

    def __getitem__(self, key) -> np.ndarray:
        if isinstance(key, (int, np.integer)):
            return self[int(key) : int(key) + 1][0]
        if not isinstance(key, slice) or key.step not in (None, 1):
            raise TypeError("SyntheticCderi supports contiguous row slices only")
        lo, hi, _ = key.indices(self.naux)
        hi = max(lo, hi)
        self.rows_generated += hi - lo
        t0 = time.perf_counter()
        out = c_binding.synth_rows(self.seed, self.n, self.scale, lo, hi - lo)
        if out is None:  # C helper not built: NumPy generator (slow, same bits)
            out = syn.synth_cderi_rows(self.seed, self.n, self.scale, np.arange(lo, hi))
        self.gen_seconds += time.perf_counter() - t0
        return out


def for_problem(p: "syn.SyntheticProblem") -> SyntheticCderi:
    return SyntheticCderi(p.seed, p.n, p.naux, p.scale)
