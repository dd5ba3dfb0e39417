"""Multi-GPU parity check (run under torchrun, one rank per GPU): aux-sharded J/K, Huzinaga SCF, mu SCF and ao2mo
must reproduce the single-GPU results computed from the full tensor.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbed_b200 import synthetic as syn  # noqa: E402
from nbed_b200.backend import NBD_HUZINAGA, NBD_MU_SHIFT, B200Context  # noqa: E402
from nbed_b200.sharding import aux_shard  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    worst = 0.0
    for n, naux, nocc, n_env, m in ((45, 37, 4, 3, 6), (300, 41, 5, 12, 10)):
        p = syn.make_problem(n=n, naux=naux, nocc=nocc, n_env=n_env, seed=0, scale=4.0 / np.sqrt(n * naux))
        b = p.cderi()
        rng = np.random.default_rng(1)
        orbs = [rng.normal(size=(n, nocc)) / np.sqrt(n), rng.normal(size=(n, nocc + 1)) / np.sqrt(n)]
        mos = syn.random_orthonormal_mos(p.ovlp, m, 0)
        import scipy.linalg

        _, c = scipy.linalg.eigh(p.hcore, p.ovlp)
        dm0 = np.array([c[:, :nocc] @ c[:, :nocc].T] * 2)

        def run(ctx):
            out = {}
            out["jk"] = np.concatenate(ctx.jk_orbitals(orbs))
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_HUZINAGA)
            c1, e1, d1, h1, info = ctx.huzinaga_scf(25, 1e-8, 1e-6, True)
            out["huz"] = np.concatenate([d1.ravel(), e1.ravel(), info["trace"].ravel()])
            ctx.scf_setup(p.nelec, p.ovlp, p.hcore, p.v_emb, p.dm_enviro, NBD_MU_SHIFT, 1e5)
            c2, e2, occ2, d2, v2, info2 = ctx.mu_scf(25, 1e-8, 0.0, dm0)
            out["mu"] = np.concatenate([d2.ravel(), info2["trace"].ravel()])
            out["ao2mo"] = ctx.ao2mo(mos[0], mos[1])
            return out

        ctx = B200Context(local)
        ctx.comm_init_from_torch()
        lo, hi = aux_shard(naux, rank, world)
        ctx.load_cderi(b[lo:hi])
        sharded = run(ctx)
        ctx.close()
        if rank == 0:
            ref_ctx = B200Context(local)  # no communicator: plain single-GPU run on the full tensor
            ref_ctx.load_cderi(b)
            ref = run(ref_ctx)
            ref_ctx.close()
            for k in ref:
                d = float(np.abs(sharded[k] - ref[k]).max()) if sharded[k].shape == ref[k].shape else float("inf")
                worst = max(worst, d)
                print(f"n={n} world={world} {k:6s} max|sharded - single| = {d:.2e}", flush=True)
        dist.barrier()
    if rank == 0:
        print("MGPU_CHECK", "OK" if worst < 1e-8 else "FAIL", worst, flush=True)
    dist.destroy_process_group()
    return 0 if worst < 1e-8 else 1


if __name__ == "__main__":
    sys.exit(main())
