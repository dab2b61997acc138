"""Multi-GPU host logic of the path: chains are sharded over ranks (one process per GPU), every rank owns a
contiguous block of GLOBAL chain ids (so the Philox streams -- keyed by global chain id -- and therefore all results
are independent of the number of ranks), and the only exchange is the acceptance statistic.

Chains never interact: there is no data-path collective.  `allreduce_acc` is the one
all-reduce(sum) the reference's `acc += 1` bookkeeping turns into (test/partialbridgenuH.jl:189).
"""
from __future__ import annotations

from typing import Tuple


def shard_chains(total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global chain id, number of chains) of `rank`: contiguous blocks, sizes differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_acc(acc: int, device=None, group=None) -> int:
    """Sum of the per-rank acceptance counters (int64).  Uses torch.distributed if it is initialised
    (NCCL on GPUs, gloo on CPU); returns `acc` unchanged in a single-process run."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return int(acc)
    t = torch.tensor([int(acc)], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())
