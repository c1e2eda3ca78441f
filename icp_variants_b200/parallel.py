"""Host-side plumbing for several GPUs (one process per GPU, torch.distributed): how the registration work is
partitioned.  No kernel lives here.

  * a SEQUENCE of independent scan pairs (alignETH's loop, main.cpp:411-498): pairs are dealt to ranks -- statically
    (`shard_pairs`) or drawn from a shared ticket counter (`PairTickets`) -- every rank runs its own queue on its own
    contexts; no collective on the data path;
  * ONE very large pair: every rank holds the whole target and a contiguous shard of the source; per iteration
    the ranks' partial normal-equation rows (<= 28 doubles) are summed and every rank solves the identical system.
    Two transports: (a) `attach_peers` + the ordinary `estimate_pose` -- the exchange runs INSIDE the reduction
    kernel over NVLink peer memory (icp_gpu_peer_*; the loop stays one CUDA graph, nothing returns to the host
    between iterations); (b) `register_sharded` -- one NCCL / gloo all-reduce per iteration through the host
    (icp_gpu_iteration_local / _apply), kept as the portable baseline the fused path is measured against.
"""
from __future__ import annotations

import numpy as np


def shard_pairs(n_pairs: int, world: int, rank: int) -> list[int]:
    """Round-robin deal of pair indices: rank r gets r, r+world, ...  (44 pairs on 8 GPUs: 6,6,6,6,5,5,5,5)."""
    return list(range(rank, n_pairs, world))


class PairTickets:
    """Dynamic deal of the pair queue: one shared counter in the process group's key-value store (the rendezvous store every
    torch.distributed job already has; `add` is an atomic fetch-and-add served by rank 0's store thread, ~0.1 ms per draw against
    milliseconds per registration).  Every rank draws the index of its next pair when one of its contexts becomes free, so a rank
    that got cheap pairs takes more of them: the pairs of a sequence differ in cost by up to 5 x (share of far queries), and with 5-6
    pairs per rank the static round-robin deal (`shard_pairs`) leaves the slowest rank 40 % behind the mean.  No collective on the
    data path -- the counter is control plane.  Every rank must be able to read every pair (a shared file system, or -- the bench --
    all scans in host memory); `key` must be the same on all ranks and fresh for every pass over the queue."""

    def __init__(self, n_pairs: int, key: str = "icp_pair_tickets", store=None):
        if store is None:
            import torch.distributed as dist
            get = getattr(dist.distributed_c10d, "_get_default_store", None)
            if get is None:
                raise RuntimeError("PairTickets: this torch has no default-store accessor; pass store= (any c10d Store with add())")
            store = get()
        self.n_pairs, self.key, self.store = int(n_pairs), str(key), store
        self.drawn: list[int] = []

    def next(self):
        """Index of the next unclaimed pair, or None when the queue is empty."""
        k = int(self.store.add(self.key, 1)) - 1
        if k >= self.n_pairs:
            return None
        self.drawn.append(k)
        return k


def shard_points(n_points: int, world: int, rank: int) -> slice:
    """Contiguous, near-equal shards of the source points."""
    base, rem = divmod(n_points, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard_points_interleaved(n_points: int, world: int, rank: int) -> slice:
    """Every world-th point, starting at `rank`: each shard is a uniform subsample of the scan, so the ranks' iterations cost the same.
    The exchange synchronises the ranks once per iteration -- with contiguous shards (different parts of the scene, different shares of
    far queries) every iteration costs the slower rank's time and the skew adds up (measured on the 3 M-point pair, 2 GPUs: 62 us per
    iteration of waiting, profiles/r2_sharded_detail_n2.json)."""
    return slice(rank, n_points, world)


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


class _DeviceRow:
    """Zero-copy torch view of the context's partial-sum row (icp_gpu_iteration_local_dev) through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def register_sharded_on_stream(ctx, n_iterations: int, init_pose=None) -> np.ndarray:
    """The same point-sharded registration with the all-reduce as a collective ON THE STREAM (ncclAllReduce of the row where it
    lies in device memory): local rows -> all-reduce -> solve are enqueued back to back, nothing returns to the host between
    iterations.  `ctx` must run on torch's current stream (ctx.set_stream).  The baseline the fused peer-memory exchange is
    measured against."""
    import torch
    import torch.distributed as dist
    pose = np.eye(4, dtype=np.float32) if init_pose is None else init_pose
    ctx.iteration_begin(pose)
    views = {}
    for _ in range(n_iterations):
        for phase in range(ctx.iteration_phases()):
            ptr, n = ctx.iteration_local_dev(phase)
            if (ptr, n) not in views:
                views[(ptr, n)] = torch.as_tensor(_DeviceRow(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(views[(ptr, n)], op=dist.ReduceOp.SUM)
            ctx.iteration_apply_dev(phase)
    return ctx.iteration_end()


def gather_peer_handles(handle: bytes) -> list[bytes]:
    """Every rank's mailbox handle in rank order (the collective that doubles as the export -> attach barrier)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(handle))
    return [bytes(h) for h in out]


def attach_peers(ctx) -> None:
    """Collective: after it, `ctx.estimate_pose()` on every rank is ONE point-sharded registration (each rank holds
    the whole target and its shard of the source) whose per-iteration all-reduce is fused into the reduction kernel."""
    import torch.distributed as dist
    handles = gather_peer_handles(ctx.peer_export())
    ctx.peer_attach(dist.get_rank(), dist.get_world_size(), handles)
    dist.barrier()      # nobody starts a registration before every rank has opened every mailbox
