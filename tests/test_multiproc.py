"""CPU tests of the host-side multi-rank logic (world_size 2, gloo): shard geometry, the
rank-ordered reduction of per-locus held-out sums, and the handle all-gather plumbing that
bench.py / SNPSamplingE use.  No GPU: the engine is replaced by a stub that returns this
rank's share of a known per-locus table."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def shard_range(n, rank, nranks):
    """Same arithmetic as SNPSamplingE.__init__ / ts_driver.hpp::create_engines."""
    per = (n + nranks - 1) // nranks
    per = (per + 3) // 4 * 4
    begin = min(rank * per, n)
    return begin, min(per, n - begin)


@pytest.mark.parametrize("n,nranks", [(200, 2), (1501, 2), (1_000_000, 8), (10, 4), (125_001, 8)])
def test_shard_geometry(n, nranks):
    covered = 0
    for r in range(nranks):
        b, m = shard_range(n, r, nranks)
        assert (b % 4 == 0 or m == 0) and b == covered and m >= 0   # an empty tail shard is refused by ts_create
        covered += m
    assert covered == n


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from terastructure_b200.snpsamplinge import SNPSamplingE, Env

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # a driver object without an engine: exercise _reduce_ll and _gather only
    s = SNPSamplingE.__new__(SNPSamplingE)
    s.rank, s.nranks, s._allgather = rank, world, allgather
    rs = np.random.RandomState(7)
    table = rs.normal(size=(world, 13))            # per-rank, per-locus partial sums
    counts = [11, 17]
    total, cnt = s._reduce_ll(table[rank].copy(), counts[rank])
    gathered = s._gather(np.full((3, 2), float(rank)))
    handles = allgather(bytes([rank]) * 64)          # comm_export -> comm_connect plumbing
    q.put((rank, total, cnt, gathered.tolist(), [h[0] for h in handles]))
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def test_rank_ordered_reduction_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    table = np.random.RandomState(7).normal(size=(world, 13))
    per = table[0] + table[1]                      # ranks added in rank order ...
    want = 0.0
    for v in per:                                  # ... then loci in ascending order
        want += float(v)
    for rank, total, cnt, gathered, hs in res:
        assert total == want and cnt == 28         # bit-identical on every rank
        assert gathered == [[0.0, 0.0]] * 3 + [[1.0, 1.0]] * 3
        assert hs == [0, 1]
