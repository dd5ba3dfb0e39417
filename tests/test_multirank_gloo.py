"""CPU, world_size 2 over gloo: the aux-index sharding contract of the multi-GPU path.

Each rank contracts only its own block of auxiliary rows; ONE all-reduce of [J | K_a | K_b] (and of the three
unique MO-integral blocks for ao2mo) reproduces the single-rank result.  The CUDA library does exactly this with
NCCL (nbd_comm_init + the all_reduce in jk / ao2mo); here the partial results come from the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nbed_b200 import synthetic as syn
from nbed_b200.sharding import aux_shard
from oracle import nbed_restatement as nr
from oracle import pyscf_restatement as ps


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = syn.make_problem(n=20, naux=37, nocc=3, n_env=2, seed=2, scale=0.05)
        lo, hi = aux_shard(p.naux, rank, world)
        b_local = p.cderi_rows(np.arange(lo, hi))
        rng = np.random.default_rng(0)
        orbs = [rng.normal(size=(p.n, 3)), rng.normal(size=(p.n, 2))]
        vj, vk = ps.df_get_jk_occ(b_local, orbs)
        buf = torch.from_numpy(np.concatenate([vj.sum(0, keepdims=True), vk]).copy())  # [J | K_a | K_b]
        dist.all_reduce(buf)
        mos = syn.random_orthonormal_mos(p.ovlp, 4, 1)
        eri = torch.from_numpy(nr.two_body_integrals(b_local, mos, restricted=False).copy())
        dist.all_reduce(eri)
        if rank == 0:
            ret["jk"] = buf.numpy()
            ret["eri"] = eri.numpy()
            ret["shards"] = [aux_shard(p.naux, r, world) for r in range(world)]
    finally:
        dist.destroy_process_group()


def test_aux_sharded_partials_allreduce_to_full_result():
    world = 2
    mgr = mp.get_context("spawn").Manager()  # (fork of a multi-threaded pytest process can deadlock)
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    p = syn.make_problem(n=20, naux=37, nocc=3, n_env=2, seed=2, scale=0.05)
    b = p.cderi()
    rng = np.random.default_rng(0)
    orbs = [rng.normal(size=(p.n, 3)), rng.normal(size=(p.n, 2))]
    vj, vk = ps.df_get_jk_occ(b, orbs)
    want = np.concatenate([vj.sum(0, keepdims=True), vk])
    assert np.abs(ret["jk"] - want).max() < 1e-12
    mos = syn.random_orthonormal_mos(p.ovlp, 4, 1)
    assert np.abs(ret["eri"] - nr.two_body_integrals(b, mos, restricted=False)).max() < 1e-12
    assert ret["shards"] == [(0, 18), (18, 37)]
