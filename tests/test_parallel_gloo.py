"""world_size-2 gloo tests (CPU) of the N>1 host logic: unit sharding, the rank-count independence of
the Philox sample streams (key = global sample index), and the result collectives."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tensortrains.jl_b200", "ttb200")):
        sys.path.insert(0, p)
    import parallel as P                      # the package __init__ needs the CUDA .so; the helper does not
    from oracle import tt_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # sample streams: shard global indices, draw with the oracle + Philox keyed by GLOBAL index
        rng = np.random.Generator(np.random.Philox(9))
        A = O.rand_tt(rng, [1, 3, 4, 3, 1], 3)
        O.normalize(A)
        rz = O.accumulate_R(A)
        n, seed = 64, 11
        lo, hi = P.shard_range(n, rank, world)
        hist = np.zeros((4, 3), dtype=np.int64)
        for i in range(lo, hi):
            x, _ = O.sample(lambda t: float(O.philox_uniform(seed, i, t)), A, rz=rz)
            for t, xt in enumerate(x):
                hist[t, xt] += 1
        tot = P.allreduce_sum(hist)
        # batch of trains: per-rank results gathered in rank order
        blo, bhi = P.shard_range(6, rank, world)
        mine = np.array([[float(b), float(b) ** 2] for b in range(blo, bhi)])
        allr = P.allgather(mine)
        tmax = P.max_over_ranks(1.0 + rank)
        q.put((rank, tot, allr, tmax, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    sys.path.insert(0, os.path.join(ROOT, "tensortrains.jl_b200", "ttb200"))
    import parallel as P
    for n in (0, 1, 7, 4096, 10 ** 8):
        for w in (1, 2, 3, 8):
            r = [P.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharding_and_collectives():
    from oracle import tt_oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference for the same global sample indices
    rng = np.random.Generator(np.random.Philox(9))
    A = O.rand_tt(rng, [1, 3, 4, 3, 1], 3)
    O.normalize(A)
    rz = O.accumulate_R(A)
    hist = np.zeros((4, 3), dtype=np.int64)
    for i in range(64):
        x, _ = O.sample(lambda t: float(O.philox_uniform(11, i, t)), A, rz=rz)
        for t, xt in enumerate(x):
            hist[t, xt] += 1
    for rank, tot, allr, tmax, rng_ in res:
        assert np.array_equal(tot, hist)
        assert np.array_equal(allr, np.array([[float(b), float(b) ** 2] for b in range(6)]))
        assert tmax == 2.0
    assert res[0][4] == (0, 32) and res[1][4] == (32, 64)
