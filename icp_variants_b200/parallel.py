"""Host-side plumbing for several GPUs (one process per GPU, torch.distributed): how the registration work is
partitioned.  No kernel lives here.

  * a SEQUENCE of independent scan pairs (alignETH's loop, main.cpp:411-498): pairs are dealt to ranks, every
    rank runs its own queue on its own context -- no collective on the data path;
  * ONE very large pair: every rank holds the whole target and a contiguous shard of the source; per iteration
    the ranks' partial normal-equation rows (<= 28 doubles) are summed with one all-reduce and every rank solves
    the identical system (icp_gpu_iteration_local / _apply).
"""
from __future__ import annotations

import numpy as np


def shard_pairs(n_pairs: int, world: int, rank: int) -> list[int]:
    """Round-robin deal of pair indices: rank r gets r, r+world, ...  (44 pairs on 8 GPUs: 6,6,6,6,5,5,5,5)."""
    return list(range(rank, n_pairs, world))


def shard_points(n_points: int, world: int, rank: int) -> slice:
    """Contiguous, near-equal shards of the source points."""
    base, rem = divmod(n_points, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def allreduce_sum(row: np.ndarray) -> np.ndarray:
    """Sum of one partial row over all ranks (gloo: CPU tensor; nccl: staged through the rank's GPU)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(row, dtype=np.float64).copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def register_sharded(ctx, n_iterations: int, init_pose=None) -> np.ndarray:
    """Point-sharded registration of one pair: `ctx` already holds the whole target and this rank's source
    shard.  Every iteration: local partial rows -> all-reduce -> identical solve on every rank."""
    pose = np.eye(4, dtype=np.float32) if init_pose is None else init_pose
    ctx.iteration_begin(pose)
    for _ in range(n_iterations):
        for phase in range(ctx.iteration_phases()):
            ctx.iteration_apply(phase, allreduce_sum(ctx.iteration_local(phase)))
    return ctx.iteration_end()
