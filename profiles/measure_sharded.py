"""Config 5b in detail: one very large pair (default 1720 sweeps x 1744 beams ~ 3 M points) registered (a) by one GPU alone, (b) point-
sharded over all ranks with ncclAllReduce on the stream, (c) with the exchange fused into the reduction kernel over peer memory -- graph
replay and launch by launch -- plus the timeline of the LAST reduction launch of the fused form on every rank (%globaltimer marks,
ICP_GPU_REDUCE_PROFILE=1).  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/measure_sharded.py
Rank 0 prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from icp_variants_b200 import capi, parallel  # noqa: E402


def main():
    sweeps = int(os.environ.get("SWEEPS", "1720")); beams = int(os.environ.get("BEAMS", "1744")); iters = int(os.environ.get("ITERS", "30"))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
    ctx = capi.Context(local); ctx.set_stream(stream.cuda_stream)
    src, tgt = bench.make_pair_device_normals(ctx, 0, sweeps, beams)
    cfg = capi.default_config()
    cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, iters, 10.0, 2, 0
    ctx.set_config(cfg)
    ctx.set_target(tgt.points, tgt.normals, tgt.colors)

    def timed(fn, reps=5):
        best, res = None, None
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); res = fn(); e1.record(stream); e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best, res

    out = {"world": world, "n_points": len(src), "iterations": iters, "shards": os.environ.get("SHARD", "interleaved")}
    if rank == 0:
        ctx.set_source(src.points, src.normals, src.colors)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.estimate_pose(want_history=False); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["single_gpu_ms"] = min(ts)
    dist.barrier()
    sl = parallel.shard_points_interleaved(len(src), world, rank) if os.environ.get("SHARD", "interleaved") == "interleaved" else parallel.shard_points(len(src), world, rank)
    ctx.set_source(src.points[sl], src.normals[sl], src.colors[sl])
    # this rank's shard alone (no exchange at all): the compute both sharded forms contain
    out_alone, _ = timed(lambda: ctx.estimate_pose(want_history=False))
    ms_nccl, _ = timed(lambda: parallel.register_sharded_on_stream(ctx, iters))
    parallel.attach_peers(ctx)
    ms_fused, _ = timed(lambda: ctx.estimate_pose(want_history=False))
    cfg.use_graph = 0; ctx.set_config(cfg)
    ms_fused_nograph, _ = timed(lambda: ctx.estimate_pose(want_history=False))
    cfg.use_graph = 1; ctx.set_config(cfg)
    ctx.estimate_pose(want_history=False)
    prof = [int(v) for v in ctx.stats().reduce_profile_ns]
    ctx.peer_detach()
    profs = [None] * world
    dist.all_gather_object(profs, prof)
    if rank == 0:
        out.update(shard_alone_ms_max_over_ranks=out_alone, nccl_allreduce_on_stream_ms=ms_nccl, fused_peer_memory_ms=ms_fused, fused_launch_by_launch_ms=ms_fused_nograph,
                   last_reduction_timeline_us_per_rank=[{"loop": (p[2] - p[1]) / 1e3, "sum_and_exchange": (p[3] - p[2]) / 1e3, "solve": (p[4] - p[3]) / 1e3,
                                                         "first_block_to_last_block_start": (p[1] - p[0]) / 1e3} for p in profs] if os.environ.get("ICP_GPU_REDUCE_PROFILE") else None)
        print(json.dumps(out))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
