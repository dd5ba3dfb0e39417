"""Shared by make_golden_fullsize.py (dev container) and tests/test_gpu_fullsize_golden.py (GPU box): how a full-size
result is condensed into a fixture small enough to commit, and how a candidate is compared with it.

A matrix M (.., n, n) is stored as every ``ROW_STRIDE``-th row in full (element-wise parity on ~3 % of the entries),
its Frobenius norm, its plain sum and a weighted checksum sum_ij w_i M_ij w_j with fixed pseudo-random weights (any
localised error outside the sampled rows moves the checksum).  A large flat tensor is stored as ``NSAMPLE`` hash-chosen
entries plus per-block sums and sums of squares.
"""
import numpy as np

from nbed_b200 import synthetic as syn

ROW_STRIDE = 32
NSAMPLE = 16384


def weights(n: int) -> np.ndarray:
    return syn.hash_uniform(977, np.arange(n, dtype=np.uint64))


def digest_matrix(m: np.ndarray) -> dict:
    m = np.asarray(m, dtype=np.float64)
    w = weights(m.shape[-1])
    return {
        "rows": np.ascontiguousarray(m[..., ::ROW_STRIDE, :]),
        "fro": np.sqrt((m * m).sum(axis=(-2, -1))),
        "sum": m.sum(axis=(-2, -1)),
        "wsum": np.einsum("i,...ij,j->...", w, m, w),
    }


def check_matrix(got: np.ndarray, gold: dict, prefix: str, tol: float, what: str):
    d = digest_matrix(got)
    scale = max(1.0, float(np.abs(gold[prefix + "_rows"]).max()))
    err = float(np.abs(d["rows"] - gold[prefix + "_rows"]).max())
    assert err < tol * scale, f"{what}: sampled rows differ by {err:.3e} (tol {tol * scale:.1e})"
    n = got.shape[-1]
    # the aggregate checks bound the error of the un-sampled entries: |d fro| <= ||dM||_F, |d wsum| <= ||dM||_F |w|^2
    for key, bound in (("fro", n), ("sum", n), ("wsum", n)):
        e = float(np.abs(d[key] - gold[f"{prefix}_{key}"]).max())
        assert e < tol * scale * bound, f"{what}: {key} differs by {e:.3e}"
    return err


def sample_indices(size: int) -> np.ndarray:
    u = syn.hash_uniform(1234, np.arange(NSAMPLE, dtype=np.uint64))
    return np.unique(((u + 1.0) * 0.5 * size).astype(np.int64).clip(0, size - 1))


def digest_tensor(t: np.ndarray) -> dict:
    t = np.asarray(t, dtype=np.float64)
    flat = t.reshape(t.shape[0], -1)
    idx = sample_indices(flat.shape[1])
    return {"samples": flat[:, idx].copy(), "sum": flat.sum(axis=1), "sumsq": (flat * flat).sum(axis=1),
            "absmax": np.abs(flat).max(axis=1)}


def check_tensor(got: np.ndarray, gold: dict, prefix: str, tol: float, what: str):
    d = digest_tensor(got)
    scale = max(1.0, float(gold[prefix + "_absmax"].max()))
    err = float(np.abs(d["samples"] - gold[prefix + "_samples"]).max())
    assert err < tol * scale, f"{what}: sampled entries differ by {err:.3e}"
    cnt = got.reshape(got.shape[0], -1).shape[1]
    assert float(np.abs(d["sum"] - gold[prefix + "_sum"]).max()) < tol * scale * np.sqrt(cnt) * 8, f"{what}: sum"
    rel = float(np.abs(d["sumsq"] / gold[prefix + "_sumsq"] - 1.0).max())
    assert rel < 1e-9, f"{what}: sum of squares differs by {rel:.3e} relative"
    return err
