"""Where one launch of reduce_kernel<point-to-plane, fused> spends its time on the benchmark pair: device timestamps
(%globaltimer) written by the kernel itself under ICP_GPU_REDUCE_PROFILE=1 (icp_gpu_stats.reduce_profile_ns), for the
last iteration of a 30-iteration registration.  Prints one JSON object (microseconds)."""
import json
import os
import sys

os.environ["ICP_GPU_REDUCE_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from icp_variants_b200 import capi  # noqa: E402


def main():
    sweeps, beams = int(os.environ.get("SWEEPS", "344")), int(os.environ.get("BEAMS", "1077"))     # 1720 x 1744: the 3 M-point pair
    cfg = capi.default_config()
    cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, 30, 10.0, 2, 0
    with capi.Context(0) as ctx:
        src, tgt = bench.make_pair_device_normals(ctx, 0, sweeps, beams) if sweeps * beams > 10 ** 6 else bench.make_pair(0, sweeps, beams)
        out = {"n_points": len(src), "algorithmic_bytes_per_launch": 48 * len(src), "runs": []}
        ctx.set_config(cfg)
        ctx.set_target(tgt.points, tgt.normals, tgt.colors)
        ctx.set_source(src.points, src.normals, src.colors)
        for _ in range(4):
            ctx.estimate_pose(want_history=False)
            t = [int(x) for x in ctx.stats().reduce_profile_ns]
            out["runs"].append({"first_block_to_last_block_start_us": (t[1] - t[0]) / 1e3, "point_loop_us": (t[2] - t[1]) / 1e3,
                                "block_reduce_ticket_final_sum_us": (t[3] - t[2]) / 1e3, "solve_pose_update_us": (t[4] - t[3]) / 1e3,
                                "first_block_start_to_pose_written_us": (t[4] - t[0]) / 1e3,
                                "algorithmic_GBps_in_kernel": 48 * len(src) / max(t[4] - t[0], 1)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
