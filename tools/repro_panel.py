import sys, numpy as np
sys.path.insert(0, '.')
from nbed_b200 import synthetic as syn
from nbed_b200.backend import B200Context
n, naux = int(sys.argv[1]), int(sys.argv[2])
nocc = [int(x) for x in sys.argv[3].split(',')]
ctx = B200Context(0)
ctx.cderi_alloc(n, naux); ctx.cderi_synth(1, 0.01, 0)
rng = np.random.default_rng(0)
orbs = [rng.normal(size=(n, o)) / np.sqrt(n) for o in nocc]
vj, vk = ctx.jk_orbitals(orbs)
print('ok', np.abs(vj).max(), np.abs(vk).max(), ctx.timers())
