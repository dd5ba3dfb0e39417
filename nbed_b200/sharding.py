"""Aux-index sharding of the 3-centre tensor across ranks: contiguous, balanced row blocks.

``cderi`` rows are independent after the Cholesky decoration, so partial J / K / MO-integral blocks computed
from disjoint row blocks simply add: one all-reduce per Fock build and one per ao2mo call."""
from __future__ import annotations


def aux_shard(naux: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of rank ``rank`` out of ``world``; sizes differ by at most one row; shards may be empty."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    return naux * rank // world, naux * (rank + 1) // world
