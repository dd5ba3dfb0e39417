"""Result-copy throughput into fresh pageable numpy arrays against the number of copy threads (development aid)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from nbed_b200 import synthetic as syn
from nbed_b200.backend import B200Context
ctx = B200Context(0)
cfg = syn.CONFIGS["C5_h2o16_def2tzvp"]; n, naux, m = cfg["n"], cfg["naux"], cfg["m"]
p = syn.make_problem(seed=1, **cfg)
ctx.cderi_alloc(n, naux); ctx.cderi_synth(1, p.scale, 0)
mo = syn.random_orthonormal_mos(p.ovlp, m, 0)
h5 = np.array([p.hcore, p.hcore])
for thr in (2, 4, 8, 12):
    ctx.set_option("copy_threads", thr)
    for name, fn in (("ao2mo", lambda: ctx.ao2mo(mo[0], mo[1])), ("build", lambda: ctx.build_hamiltonian(h5, mo[0], mo[1]))):
        fn()
        ts = []
        for _ in range(4):
            t = time.perf_counter(); out = fn(); ts.append((time.perf_counter() - t) * 1e3)
        dev = ctx.timer_ms("ao2mo_total") if name == "ao2mo" else ctx.timer_ms("build_total")
        print(f"copy_threads={thr} {name}: wall ms {[round(x, 1) for x in ts]} device {dev:.2f}", flush=True)
