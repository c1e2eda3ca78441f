"""Config 5b: one very large pair (default 1720 sweeps x 1744 beams ~ 3 M points) registered (a) by one GPU alone and
(b) point-sharded over all ranks with one all-reduce of <= 28 doubles per iteration (NCCL through the host, and fused
into the reduction kernel over peer memory).  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/measure_sharded.py
Rank 0 prints one JSON object.  ("used only where it is measured to win": this is that measurement.)"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import capi, parallel, synth  # noqa: E402


def main():
    sweeps = int(os.environ.get("SWEEPS", "1720")); beams = int(os.environ.get("BEAMS", "1744")); iters = int(os.environ.get("ITERS", "30"))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    src, tgt, _ = synth.eth_pair(seed=1234, n_sweeps=sweeps, n_beams=beams)      # every rank builds the same pair
    cfg = capi.default_config()
    cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, iters, 10.0, 2, 0
    out = {}
    ctx = capi.Context(local)
    ctx.set_config(cfg)
    ctx.set_target(tgt.points, tgt.normals, tgt.colors)
    # (a) one GPU, whole source
    if rank == 0:
        ctx.set_source(src.points, src.normals, src.colors)
        for _ in range(2):
            t0 = time.perf_counter(); pose_single, _, _ = ctx.estimate_pose(want_history=False); dt = time.perf_counter() - t0
        out["single_gpu_ms"] = dt * 1e3
    dist.barrier()
    # (b) sharded source
    sl = parallel.shard_points(len(src), world, rank)
    ctx.set_source(src.points[sl], src.normals[sl], src.colors[sl])
    for _ in range(2):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        pose_sharded = parallel.register_sharded(ctx, iters)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
    dt = parallel.max_over_ranks(dt)
    # (c) the same shards, exchange fused into the reduction kernel over NVLink peer memory (icp_gpu_peer_*): the
    # ordinary estimate_pose, one CUDA graph, no host round trip and no NCCL call per iteration
    parallel.attach_peers(ctx)
    for _ in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        pose_fused, _, _ = ctx.estimate_pose(want_history=False)
        dtf = time.perf_counter() - t0
    dtf = parallel.max_over_ranks(dtf)
    poses = [None] * world
    dist.all_gather_object(poses, pose_fused.tobytes())
    ctx.peer_detach()
    if rank == 0:
        out.update(sharded_nccl_host_staged_ms=dt * 1e3, sharded_fused_peer_memory_ms=dtf * 1e3, world=world, n_points=len(src), iterations=iters,
                   max_abs_pose_diff_nccl=float(np.abs(pose_sharded - pose_single).max()),
                   max_abs_pose_diff_fused=float(np.abs(pose_fused - pose_single).max()),
                   fused_pose_identical_on_all_ranks=bool(all(p == poses[0] for p in poses)),
                   note="nccl: the 28-double row goes D2H, NCCL all-reduce, H2D every iteration; fused: the last block of the reduction "
                        "kernel stores its row into the peers' mailboxes (NVLink) and sums what it receives")
        print(json.dumps(out))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
