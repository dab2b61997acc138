"""Multi-GPU host logic of the path: chains are sharded over ranks (one process per GPU), every rank owns a
contiguous block of GLOBAL chain ids (so the Philox streams -- keyed by global chain id -- and therefore all results
are independent of the number of ranks), and the only exchange is the acceptance statistic.

Chains never interact: there is no data-path collective.  The one all-reduce(sum) the
reference's `acc += 1` bookkeeping turns into (test/partialbridgenuH.jl:189) is issued by the LIBRARY
(bb_allreduce_acc: NCCL on the communicator's own stream, behind an event -- never on the compute stream);
`Communicator` wraps it.  torch.distributed (or any other launcher) is only used to hand the 128-byte NCCL id from
rank 0 to the other ranks.  `allreduce_acc` is the host-side equivalent for CPU-only runs (gloo).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np


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


def exchange_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 calls `make_id()` (-> 128 bytes); every rank returns the same bytes.  Transport: torch.distributed's
    default group (any backend) when it is initialised; a single rank needs none."""
    if world == 1:
        return make_id()
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("exchange_unique_id: initialise torch.distributed first (or ship the id yourself)")
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return bytes(box[0])


class Communicator:
    """bb_comm: NCCL communicator + side stream for the acceptance-counter all-reduce of one context."""

    def __init__(self, ctx, rank: int = 0, world: int = 1, unique_id: Optional[bytes] = None):
        from . import _cabi as K
        self._K, self.ctx, self.rank, self.world = K, ctx, rank, world
        if unique_id is None:
            unique_id = exchange_unique_id(self.make_unique_id, rank, world)
        if len(unique_id) != K.NCCL_ID_BYTES:
            raise ValueError("unique_id: 128 bytes")
        buf = (C.c_uint8 * K.NCCL_ID_BYTES).from_buffer_copy(unique_id)
        h = C.c_void_p()
        K.check(K.lib.bb_comm_create(ctx.h, world, rank, buf, C.byref(h)))
        self.h = h

    @staticmethod
    def make_unique_id() -> bytes:
        from . import _cabi as K
        buf = (C.c_uint8 * K.NCCL_ID_BYTES)()
        K.check(K.lib.bb_comm_unique_id(buf))
        return bytes(buf)

    def allreduce_acc_(self, ens, theta: bool = False):
        """Start the all-reduce of the ensemble's acceptance counter as of the calls issued so far (asynchronous)."""
        f = self._K.lib.bb_allreduce_theta_acc if theta else self._K.lib.bb_allreduce_acc
        self._K.check(f(ens.h, self.h))

    @property
    def acc(self) -> int:
        """Global sum delivered by the most recent allreduce_acc_ (waits for that all-reduce only)."""
        v = C.c_int64(0)
        self._K.check(self._K.lib.bb_comm_get_acc(self.h, C.byref(v)))
        return v.value

    def synchronize(self):
        self._K.check(self._K.lib.bb_comm_synchronize(self.h))

    def close(self):
        if self.h:
            self._K.lib.bb_comm_destroy(self.h)
            self.h = None
