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
