"""GPU: bench.py prints exactly one JSON line with the contract's keys (small workload so it runs in seconds)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "C3", "--steps", "6", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["metric"] == "embedded_scf_iterations_per_s" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 6 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-6 * 1e3
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    assert d["ao2mo"]["gbs"] > 0 and d["hamiltonian_build"]["wall_ms"] > 0
    assert len(d["checksum"]["energy_last_step"]) == 2 and "FULL iterations" in d["cpu_baseline"]["sample"]
    assert {"n", "naux", "nocc_per_spin", "n_env", "sharding"} <= set(d["config"])
