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


def _oracle_block(A, lo, hi, md):
    """the pairs with first site in [lo, hi) from the CPU oracle, t-major, flattened column-major"""
    from oracle import tt_oracle as O
    m2 = O.twovar_marginals(A, maxdist=md)
    L = len(A)
    rows = [m2[(t, u)].reshape(-1, order="F") for t in range(lo, hi) for u in range(t + 1, min(L - 1, t + md) + 1)]
    Q = int(np.prod(A[0].shape[2:]))
    return np.array(rows).reshape(len(rows), Q * Q)


def _worker_pairs(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tensortrains.jl_b200", "ttb200")):
        sys.path.insert(0, p)
    import parallel as P
    from oracle import tt_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.Generator(np.random.Philox(19))
        A = O.rand_tt(rng, [1, 3, 4, 2, 3, 2, 1], 2, 2)
        res = {}
        L = len(A)
        for md in (None, 2):
            mdv = L - 1 if md is None else md
            lo, hi = P.pair_ranges(L, mdv, world)[rank]
            full = P.gather_pair_blocks(L, A[0].shape[2:], mdv, _oracle_block(A, lo, hi, mdv))
            res[md] = {k: v.copy() for k, v in full.items()}
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_pair_ranges_partition():
    sys.path.insert(0, os.path.join(ROOT, "tensortrains.jl_b200", "ttb200"))
    import parallel as P
    for L, md, w in ((128, 127, 8), (64, 63, 3), (10, 3, 4), (5, 4, 8), (2, 1, 2), (7, 6, 1)):
        r = P.pair_ranges(L, md, w)
        assert len(r) == w and r[0][0] == 0 and r[-1][1] == L - 1
        assert all(r[k][1] == r[k + 1][0] and r[k][0] <= r[k][1] for k in range(w - 1))
        cnt = [sum(min(L - 1, t + md) - t for t in range(a, b)) for a, b in r]
        per_t = max(min(L - 1, t + md) - t for t in range(L - 1))
        assert max(cnt) - min(c for c in cnt) <= 2 * per_t or w > L - 1      # balanced up to one site's pairs


@pytest.mark.timeout(300)
def test_two_rank_sharded_pair_marginals():
    """the pair marginals of ONE train sharded over two ranks by first site, all-gathered (ragged blocks)"""
    from oracle import tt_oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_pairs, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.Generator(np.random.Philox(19))
    A = O.rand_tt(rng, [1, 3, 4, 2, 3, 2, 1], 2, 2)
    for md in (None, 2):
        ref = O.twovar_marginals(A, maxdist=md)
        for rank, got in res:
            assert set(got[md]) == set(ref)
            for k in ref:
                assert np.array_equal(got[md][k], ref[k])
