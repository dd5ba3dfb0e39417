"""CPU: the C-ABI library loads and exports every declared symbol; host-side helpers; the bench reference arm."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    from nbed_b200 import _lib, build

    build.build_library()  # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    lib = _lib.load()
    assert lib.nbd_version() >= 100
    header = open(os.path.join(ROOT, "include", "nbed_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(nbd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nbed_b200 import B200Context, NbdError

    with pytest.raises(NbdError):
        B200Context(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nbed_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_aux_shards_tile_the_index():
    from nbed_b200.sharding import aux_shard

    for naux in (0, 1, 7, 4128, 4129):
        for world in (1, 2, 3, 8):
            edges = [aux_shard(naux, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == naux
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        aux_shard(10, 2, 2)


def test_synthetic_tensor_is_counter_based():
    from nbed_b200 import synthetic as syn

    full = syn.synth_cderi_rows(3, 9, 0.5, np.arange(12))
    part = syn.synth_cderi_rows(3, 9, 0.5, np.array([4, 11]))
    assert np.array_equal(part, full[[4, 11]])
    assert full.min() >= -0.5 and full.max() < 0.5 and abs(full.mean()) < 0.05
    # known answer of the SplitMix64 stream (guards the device generator's bit-compatibility contract)
    v = syn.hash_uniform(5, np.array([0, 1, 2 ** 40], dtype=np.uint64))
    assert np.allclose(v, [0.9413689876, -0.7110414177, 0.7827159049], atol=1e-9), v
    p = syn.make_problem(n=12, naux=20, nocc=3, n_env=2, seed=4)
    assert np.abs(p.c_env[0].T @ p.ovlp @ p.c_env[0] - np.eye(2)).max() < 1e-12
    assert p.cderi().shape == (20, 78)


def test_streamed_tensor_matches_the_numpy_generator_bit_for_bit():
    """oracle/streamed.py (C generator behind a row-sliceable object) == synthetic.synth_cderi_rows, and the oracle's
    J/K over the streamed object == over the materialised array: what the full-size goldens rest on."""
    from nbed_b200 import synthetic as syn
    from oracle import pyscf_restatement as ps
    from oracle import streamed

    p = syn.make_problem(n=37, naux=301, nocc=3, n_env=2, seed=6)
    st = streamed.for_problem(p)
    full = p.cderi()
    assert st.shape == full.shape and len(st) == 301
    assert np.array_equal(st[0:301], full) and np.array_equal(st[250:999], full[250:]) and np.array_equal(st[7], full[7])
    rng = np.random.default_rng(0)
    orbs = [rng.normal(size=(37, 3)), rng.normal(size=(37, 2))]
    j0, k0 = ps.df_get_jk_occ(full, orbs)
    j1, k1 = ps.df_get_jk_occ(st, orbs)
    assert np.array_equal(j0, j1) and np.array_equal(k0, k1)
    mf = ps.DFUHF(p.ovlp, p.hcore, st, p.nelec)
    assert mf.cderi is st


def test_full_size_goldens_are_committed_and_describe_the_stated_sizes():
    g4 = np.load(os.path.join(ROOT, "tests", "golden", "c4_fullsize.npz"))
    g5 = np.load(os.path.join(ROOT, "tests", "golden", "c5_fullsize.npz"))
    assert (int(g4["n"]), int(g4["naux"])) == (1376, 4128) and int(g4["cycles"]) >= 3
    assert g4["energies"].shape == (int(g4["cycles"]), 2) and g4["jk_orbitals"].shape == (2, 1376, 5)
    assert g4["oracle_vs_reference"].max() < 1e-10  # restatement == unmodified reference loop at the full size
    for name in ("vj", "vk", "dm", "huz"):
        assert g4[name + "_rows"].shape[-2:] == (43, 1376) and np.isfinite(g4[name + "_fro"]).all()
    assert (int(g5["n"]), int(g5["naux"]), int(g5["m"])) == (688, 2064, 40)
    assert g5["two_samples"].shape[0] == 4 and g5["two_samples"].shape[1] > 10000
    from nbed_b200 import synthetic as syn

    _, p = syn.bench_problem("C4_h2o32_def2tzvp")
    assert p.scale == float(g4["scale"]) and p.seed == int(g4["seed"])


def test_diis_restatement_recovers_linear_fixed_point():
    """pyscf.lib.diis.DIIS on x -> A x + b converges to the fixed point in <= dim + 2 steps."""
    from oracle.pyscf_restatement import DIIS

    rng = np.random.default_rng(0)
    a = rng.normal(size=(4, 4)) * 0.3
    b = rng.normal(size=4)
    fix = np.linalg.solve(np.eye(4) - a, b)
    d, x = DIIS(), np.zeros(4)
    for _ in range(8):
        x = d.update(a @ x + b)
    assert np.abs(x - fix).max() < 1e-10


def test_bench_reference_arm_emits_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C3",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "embedded_scf_iterations_per_s"
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert "FULL iterations" in line["cpu_baseline"]["sample"] and "extrapolated" in line["cpu_baseline"]["sample"]
    # same `config` keys as the GPU arm prints (the driver compares the two)
    assert {"workload", "n", "naux", "nocc_per_spin", "n_env", "sharding"} <= set(line["config"])
    assert line["config"]["n"] == 174 and line["config"]["naux"] == 522
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_panel_task_lists_keep_their_ring_invariants(tmp_path):
    """Host-side check of the panel kernel's per-warp event lists for every matrix size / ring depth: consecutive
    events of a warp at most S tiles apart (cyclically), arrival weights of every tile summing to the barrier count."""
    exe = tmp_path / "plan_test"
    src = os.path.join(ROOT, "tests", "cpp", "plan_test.cu")
    subprocess.run(["nvcc", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe), src], check=True, timeout=300)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout[-500:]


def test_host_linear_algebra_of_the_scf_drivers(tmp_path):
    """The library's pure-C++ host maths (Jacobi eigensolver, LU, DIIS coefficients incl. the pseudo-inverse branch of
    pyscf/lib/diis.py, the Rayleigh-Ritz step of the subspace eigensolver) against manufactured solutions."""
    exe = tmp_path / "linalg_test"
    src = os.path.join(ROOT, "tests", "cpp", "linalg_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), src], check=True, timeout=300)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "fails=0" in out.stdout, out.stdout[-800:]


def test_result_pool_recycles_page_locked_blocks(monkeypatch):
    """nbed_b200._lib.result_empty: results of 1 MiB and more live in blocks of nbd_host_alloc that return to a pool when
    the array AND every view of it have been garbage collected, and are handed out again for the next result of that
    size; smaller results are plain NumPy arrays; a failed allocation falls back to a pageable array."""
    import ctypes
    import gc

    from nbed_b200 import _lib

    libc = ctypes.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes, libc.free.argtypes = ctypes.c_void_p, [ctypes.c_size_t], [ctypes.c_void_p]

    class Fake:
        def __init__(self):
            self.allocs, self.frees, self.fail = [], [], False

        def nbd_host_alloc(self, n):
            if self.fail:
                return None
            p = libc.malloc(n)
            self.allocs.append(p)
            return p

        def nbd_host_free(self, p):
            self.frees.append(p)
            libc.free(p)

    fake = Fake()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "_POOL", {})
    monkeypatch.setattr(_lib, "_POOL_STATE", {"bytes": 0, "cap": 8 << 20})
    a = _lib.result_empty((512, 512))
    a[:] = 1.0
    addr, view = a.ctypes.data, a[:10]
    del a
    gc.collect()
    assert not any(_lib._POOL.values()) and view.sum() == 5120.0  # a view keeps the block out of the pool
    del view
    gc.collect()
    assert sum(len(v) for v in _lib._POOL.values()) == 1 and _lib._POOL_STATE["bytes"] == 512 * 512 * 8
    b = _lib.result_empty((512, 512))
    assert b.ctypes.data == addr and len(fake.allocs) == 1 and _lib._POOL_STATE["bytes"] == 0
    assert len(fake.allocs) == 1 and isinstance(_lib.result_empty((4, 4)), np.ndarray) and len(fake.allocs) == 1
    big = [_lib.result_empty((1024, 1024)) for _ in range(2)]  # 2 x 8 MiB against a cap of 8 MiB: one is freed for real
    del big, b
    gc.collect()
    assert _lib._POOL_STATE["bytes"] <= 8 << 20 and len(fake.frees) >= 1
    fake.fail = True
    c = _lib.result_empty((600, 600))
    c[:] = 2.0
    assert c.shape == (600, 600) and c.sum() == 720000.0
