"""Does the group search pay on a point shard?  One GPU, the 3 M-point target, the source = every 8th point (the interleaved
shard of one of 8 ranks, parallel.shard_points_interleaved): 30 iterations with the group search off / on."""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi
gen = capi.Context(0)
src, tgt = bench.make_pair_device_normals(gen, 0, 1720, 1744)
out = {"n_target": len(tgt), "n_shard": len(src.points[0::8])}
for gm in (0, 8):
    os.environ["ICP_GPU_GROUP_MIN"] = str(gm)
    c = capi.Context(0)
    cfg = capi.default_config(); cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, 30, 10.0, 2, 0
    c.set_config(cfg); c.set_target(tgt.points, tgt.normals, tgt.colors); c.set_source(src.points[0::8], src.normals[0::8], src.colors[0::8])
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pose, _, _ = c.estimate_pose(want_history=False); e1.record(); e1.synchronize()
        best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
    out[f"ms_group_min_{gm}"] = best; out[f"checksum_{gm}"] = float(np.abs(pose).sum())
    c.close()
print(json.dumps(out))
