"""Shape-synthetic embedded-SCF problems shared by tests, bench and the oracle.

The reference's integral source (libcint, via pyscf.gto) is a stated boundary of the hot path and is not
available in this image, so throughput and kernel parity are measured on synthetic tensors with the
shapes of BASELINE.json's configs.  The 3-centre tensor ``B[P, mu>=nu]`` is defined by a counter-based
hash (SplitMix64) of its packed index, so that the same values can be produced on the host for any
subset of auxiliary rows (oracle sample) and on the device at full size (31 GB at (H2O)32/def2-TZVP;
``nbd_synth_cderi`` in csrc/).  Everything O(n^2) comes from ``numpy.random.default_rng``.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# n = nao, A = naux (A ~ 3n ESTIMATE), o = active occupied per spin, m = active-space MOs per spin
CONFIGS = {
    "C1_h2o_sto3g": dict(n=7, naux=21, nocc=4, n_env=1, m=5),
    "C2_h2o_ccpvdz": dict(n=24, naux=72, nocc=4, n_env=1, m=12),
    "C3_ethanol_ccpvtz": dict(n=174, naux=522, nocc=9, n_env=4, m=24),
    "C4_h2o32_def2tzvp": dict(n=1376, naux=4128, nocc=5, n_env=155, m=40),
    # the same system with a 20-orbital active region per spin (the FP64-tensor-bound regime)
    "C4_h2o32_def2tzvp_o20": dict(n=1376, naux=4128, nocc=20, n_env=140, m=40),
    "C5_h2o16_def2tzvp": dict(n=688, naux=2064, nocc=5, n_env=75, m=40),
}

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash_uniform(seed: int, index: np.ndarray) -> np.ndarray:
    """Uniform in [-1, 1): bit-identical to ``synth_uniform`` in csrc/nbed_kernels.cu."""
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) << np.uint64(48)) + index.astype(np.uint64)
    z = splitmix64(key)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0


def synth_cderi_rows(seed: int, n: int, scale: float, rows: np.ndarray) -> np.ndarray:
    """Packed-lower rows ``B[P, mu(mu+1)/2+nu]`` for the auxiliary indices in ``rows``."""
    npair = n * (n + 1) // 2
    rows = np.asarray(rows, dtype=np.uint64)
    idx = rows[:, None] * np.uint64(npair) + np.arange(npair, dtype=np.uint64)[None, :]
    return hash_uniform(seed, idx) * scale


def default_scale(n: int, naux: int) -> float:
    # ||J|| ~ O(1): sigma ~ 1/sqrt(A n)  (SURVEY.md §8(d)); uniform[-1,1) has std 1/sqrt(3)
    return 0.9 / np.sqrt(float(naux) * float(n))


@dataclasses.dataclass
class SyntheticProblem:
    n: int
    naux: int
    nocc: int
    n_env: int
    seed: int
    scale: float
    ovlp: np.ndarray  # (n, n)
    hcore: np.ndarray  # (n, n)
    v_emb: np.ndarray  # (2, n, n) "embedding_potential"
    c_env: np.ndarray  # (2, n, n_env) S-orthonormal environment orbitals
    dm_enviro: np.ndarray  # (2, n, n)
    nelec: tuple

    def cderi_rows(self, rows) -> np.ndarray:
        return synth_cderi_rows(self.seed, self.n, self.scale, np.asarray(rows))

    def cderi(self) -> np.ndarray:
        return self.cderi_rows(np.arange(self.naux))


def make_problem(n: int, naux: int, nocc: int, n_env: int, seed: int = 0, scale: float | None = None,
                 spin_polarised_potential: bool = True, **_unused) -> SyntheticProblem:
    rng = np.random.default_rng(seed)
    r = rng.normal(0.0, 0.2 / np.sqrt(n), size=(n, n))
    ovlp = np.eye(n) + 0.5 * (r + r.T)
    r2 = rng.normal(0.0, 0.05, size=(n, n))
    hcore = np.diag(-np.linspace(10.0, 0.5, n)) + 0.5 * (r2 + r2.T) / np.sqrt(n / 16.0)
    w, v = np.linalg.eigh(ovlp)
    x = (v / np.sqrt(w)) @ v.T
    e, c = np.linalg.eigh(x @ hcore @ x)
    c_std = x @ c
    # environment = the n_env orbitals just above the o lowest
    c_env_a = c_std[:, nocc : nocc + n_env]
    c_env = np.array([c_env_a, c_env_a])
    dm_enviro = np.array([c_env_a @ c_env_a.T] * 2)
    r3 = rng.normal(0.0, 0.02, size=(2, n, n)) / np.sqrt(n / 16.0)
    v_emb = 0.5 * (r3 + r3.transpose(0, 2, 1))
    if not spin_polarised_potential:
        v_emb[1] = v_emb[0]
    if scale is None:
        scale = default_scale(n, naux)
    return SyntheticProblem(n, naux, nocc, n_env, seed, float(scale), ovlp, hcore, v_emb, c_env, dm_enviro,
                            (nocc, nocc))


# Coupling strength of bench.py's synthetic tensor, in multiples of 1/sqrt(naux * nao): chosen so that the embedded SCF
# needs a realistic ~10 cycles from the core guess.  The full-size golden fixtures (tests/golden/make_golden_fullsize.py)
# are generated on exactly this problem, so parity at the stated sizes is parity of the benchmarked workload.
BENCH_COUPLING = 16.0
BENCH_SEED = 1


def bench_problem(key: str, coupling: float | None = None) -> tuple[dict, SyntheticProblem]:
    cfg = dict(CONFIGS[key])
    c = BENCH_COUPLING if coupling is None else float(coupling)
    return cfg, make_problem(seed=BENCH_SEED, scale=c / np.sqrt(cfg["n"] * cfg["naux"]), **cfg)


def random_orthonormal_mos(ovlp: np.ndarray, m: int, seed: int = 0) -> np.ndarray:
    """(2, n, m) S-orthonormal coefficient blocks for the ao2mo tests."""
    rng = np.random.default_rng(seed + 77)
    n = ovlp.shape[0]
    w, v = np.linalg.eigh(ovlp)
    x = (v / np.sqrt(w)) @ v.T
    out = []
    for _ in range(2):
        q, _r = np.linalg.qr(rng.normal(size=(n, m)))
        out.append(x @ q)
    return np.array(out)
