"""One-process-per-GPU sharding helpers (torch.distributed is plumbing only).

The hot path shards along independent units -- trains of a batch, sample streams, sites of a
marginal computation -- with NO data-path collective: tensors are never
exchanged, a single train's sweep stays on one GPU.  Collectives only combine small results:
all-reduce of sample histograms / statistics, all-gather of per-train scalars.
Works with the `nccl` backend on GPUs and with `gloo` on CPU (used by the world_size-2 tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `n_units` for `rank`; slices differ in size by at most one."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_sum(arr: np.ndarray) -> np.ndarray:
    """Sum a small result array (e.g. the L x Q sample histogram) over all ranks."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allgather(arr: np.ndarray) -> np.ndarray:
    """Concatenate equally-shaped per-rank result arrays along axis 0 (per-train log Z, bonds, ...)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device(dist))
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return torch.cat(outs, dim=0).cpu().numpy()


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pair_ranges(L: int, maxdist: int, world: int):
    """Contiguous ranges [t_lo, t_hi) of first sites, one per rank, with (almost) equal numbers of pairs
    (t, u), t < u <= t + maxdist -- the third sharding axis of SURVEY.md 8e: the pair marginals of ONE train.
    The environments are recomputed on every rank (one streaming pass); only the O(L^2) pair loop is split."""
    counts = [min(L - 1, t + maxdist) - t for t in range(L - 1)]
    total = sum(counts)
    bounds, acc, r = [0], 0, 1
    for t, c in enumerate(counts):
        acc += c
        while r < world and acc * world >= total * r:
            bounds.append(t + 1)
            r += 1
    while len(bounds) < world:
        bounds.append(L - 1)
    bounds.append(L - 1)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


def allgather_ragged(arr: np.ndarray, lengths) -> np.ndarray:
    """Concatenate per-rank arrays whose axis-0 lengths differ (`lengths[k]` for rank k, known to all ranks):
    pad to the longest, all-gather, trim."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return arr
    m = max(int(n) for n in lengths)
    pad = np.zeros((m,) + arr.shape[1:], dtype=arr.dtype)
    pad[: arr.shape[0]] = arr
    full = allgather(pad[None])                     # (world, m, ...)
    return np.concatenate([full[k, : int(lengths[k])] for k in range(len(lengths))], axis=0)


def _pairs_dict(L: int, q, maxdist: int, full: np.ndarray):
    q = tuple(q)
    out, i = {}, 0
    for t in range(L - 1):
        for u in range(t + 1, min(L - 1, t + maxdist) + 1):
            out[(t, u)] = full[i].reshape(q + q, order="F")
            i += 1
    return out


def pair_block_lengths(L: int, maxdist: int, world: int):
    """Number of pairs in every rank's block of ``pair_ranges``."""
    return [sum(min(L - 1, t + maxdist) - t for t in range(a, b)) for a, b in pair_ranges(L, maxdist, world)]


def gather_pair_blocks(L: int, q, maxdist: int, mine: np.ndarray):
    """All-gather the per-rank blocks of pair marginals (rank k holds the pairs whose first site lies in
    ``pair_ranges(L, maxdist, world)[k]``, t-major, one row of Q*Q values per pair) and return the reference's
    dict ``(t, u) -> array`` shaped ``q + q``.  Host arrays in, host arrays out (any backend)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    full = allgather_ragged(np.ascontiguousarray(mine), pair_block_lengths(L, maxdist, world))
    return _pairs_dict(L, q, maxdist, full)


def allgather_device_rows(mine, lengths, host_out=None):
    """NCCL path of the ragged gather: ``mine`` is this rank's DEVICE tensor ``(max(lengths), width)`` whose first
    ``lengths[rank]`` rows are valid.  One ``all_gather_into_tensor`` between device buffers, then the valid rows
    go device -> PINNED host (``host_out``: a pinned float64 tensor ``(sum(lengths), width)``, allocated when
    absent).  No pageable staging, no host round trip before the collective.  Returns the pinned host tensor."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    full = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(full, mine.contiguous())
    tot = int(sum(int(n) for n in lengths))
    if host_out is None:
        host_out = torch.empty((tot,) + tuple(mine.shape[1:]), dtype=mine.dtype, pin_memory=True)
    off = 0
    for k, n in enumerate(lengths):
        n = int(n)
        if n:
            host_out[off:off + n].copy_(full[k, :n], non_blocking=True)
        off += n
    torch.cuda.current_stream(mine.device).synchronize()
    return host_out


def twovar_marginals_sharded(A, maxdist=None, as_dict: bool = True):
    """``twovar_marginals`` of one device train with the pairs sharded over the ranks (every rank holds the
    train): rank k computes the pairs whose first site lies in ``pair_ranges(...)[k]`` on its GPU
    (``ttb_twovar_marginals_range``) and the ragged blocks are all-gathered.  With the ``nccl`` backend the
    block never leaves the device before the collective (``allgather_device_rows``); with ``gloo`` (CPU tests)
    it travels as a host array.  ``as_dict=False`` returns the packed ``(pairs, Q*Q)`` array instead."""
    import torch.distributed as dist
    import ttb200 as T
    L = A.L
    md = L - 1 if maxdist is None else int(maxdist)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = pair_ranges(L, md, world)[rank]
    if world > 1 and dist.get_backend() == "nccl" and A.batch == 1:
        import torch
        lengths = pair_block_lengths(L, md, world)
        mine = torch.zeros((max(max(lengths), 1), A.Q * A.Q), dtype=torch.float64, device=_device(dist))
        T.twovar_marginals_block(A, lo, hi, md, out=mine)
        full = allgather_device_rows(mine, lengths).numpy()
    else:
        mine = T.twovar_marginals_block(A, lo, hi, md)[0]
        full = allgather_ragged(np.ascontiguousarray(mine), pair_block_lengths(L, md, world)) if world > 1 else mine
    return _pairs_dict(L, A.q, md, full) if as_dict else full
