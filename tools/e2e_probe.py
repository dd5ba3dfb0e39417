"""Where does the end-to-end call spend its wall time?  Runs nbed_b200.scf.huzinaga_scf's steps one by one at the C4
shape (host NumPy in / out) and prints wall clock per step next to the library's device stage timers.
usage: python tools/e2e_probe.py [--cycles 20] [--option k=v ...]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from nbed_b200.backend import B200Context, NBD_HUZINAGA  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cycles", type=int, default=20)
ap.add_argument("--workload", default="C4")
ap.add_argument("--option", action="append", default=[])
a = ap.parse_args()
key, _ = bench.WORKLOADS[a.workload]
cfg, p = bench.build_problem(key)
n, naux = cfg["n"], cfg["naux"]
ctx = B200Context(0)
for kv in a.option:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ctx.cderi_alloc(n, naux)
ctx.cderi_synth(p.seed, p.scale, 0)
for rep in range(3):
    t = [time.perf_counter()]
    ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
    t.append(time.perf_counter())
    ctx.scf_set_env_orbitals(p.c_env)
    t.append(time.perf_counter())
    c, e, dm, huz, info = ctx.huzinaga_scf(a.cycles, 0.0, 0.0, True)
    t.append(time.perf_counter())
    tm = ctx.timers()
    d = [round(1e3 * (t[i + 1] - t[i]), 2) for i in range(3)]
    print(f"rep {rep}: setup {d[0]} ms, env_orbitals {d[1]} ms, scf call {d[2]} ms, total {round(sum(d), 2)} ms; "
          f"device: scf_total {tm.get('scf_total', 0):.2f}, iter_total {tm.get('iter_total', 0):.2f}, eigh {tm.get('eigh', 0):.2f}, "
          f"orth {tm.get('orth', 0):.2f}", flush=True)
ctx.close()
