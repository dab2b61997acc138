"""Host-side multi-GPU logic on CPU: 2 gloo ranks (no GPU): shard arithmetic, the acceptance all-reduce, and the
property that makes sharding transparent -- random streams are keyed by GLOBAL chain id (checked with the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(ROOT, "bridge.jl_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    from oracle import oracle as O
    orc = O.load("ref")
    first, count = sh.shard_chains(total, rank, world)
    # every rank "accepts" chain c at iteration 0 iff logU(c) <= -0.5: depends on the global chain id only
    acc = sum(1 for c in range(first, first + count) if orc.accept_logu(7, 0, c) <= -0.5)
    tot = sh.allreduce_acc(acc)
    out.put((rank, first, count, acc, tot))
    dist.destroy_process_group()


def test_shard_chains_partition():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(ROOT, "bridge.jl_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    for total, world in ((250000, 8), (10, 3), (7, 8), (0, 2)):
        seen = []
        for r in range(world):
            f, c = sh.shard_chains(total, r, world)
            seen += list(range(f, f + c))
        assert seen == list(range(total))
    assert sh.shard_chains(250000, 3, 8) == (93750, 31250)  # BASELINE config 4: 31 250 chains per GPU on 8 GPUs
    with pytest.raises(ValueError):
        sh.shard_chains(10, 2, 2)
    assert sh.allreduce_acc(5) == 5  # no process group: identity


def test_two_rank_gloo_allreduce_matches_single_process(oracle_ref):
    import torch.multiprocessing as mp
    total, world = 101, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = sum(1 for c in range(total) if oracle_ref.accept_logu(7, 0, c) <= -0.5)
    assert [r[1] for r in res] == [0, 51] and [r[2] for r in res] == [51, 50]
    assert sum(r[3] for r in res) == want
    assert all(r[4] == want for r in res)  # every rank sees the global acceptance count


def test_streams_do_not_depend_on_sharding(oracle_ref):
    """The Philox row of (chain, segment) is (chain_offset + local index) * S + segment with chain_offset =
    shard_chains(...)[0] (bb_ens_set_chain_offset): for every world size the ranks' rows tile the single-process
    rows exactly -- same ids, each once -- so every chain draws the same noise however the chains are sharded
    (tests/test_gpu_parity.py checks the device side: chain_offset enters the kernels' counters)."""
    from bridge_jl_b200.sharding import shard_chains
    total, S = 1003, 4
    one = [(c * S + seg) for c in range(total) for seg in range(S)]
    for world in (2, 3, 8):
        rows = []
        for r in range(world):
            first, count = shard_chains(total, r, world)
            rows += [((first + loc) * S + seg) for loc in range(count) for seg in range(S)]
        assert rows == one
    # and the rows of different chains are different streams
    tt = np.linspace(0, 1, 33)
    first, count = shard_chains(total, 5, 8)
    a = oracle_ref.wiener_sample(tt, 1, 4, 9, (first + 0) * S + 2)
    b = oracle_ref.wiener_sample(tt, 1, 4, 9, shard_chains(total, 0, 1)[0] * S + first * S + 2)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, oracle_ref.wiener_sample(tt, 1, 4, 9, (first + 1) * S + 2))
