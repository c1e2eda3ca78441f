"""knn_group_kernel: what it takes and what it leaves (instrumented registration: windows seen, groups started / finished, members
finished, leaves listed, leaf scans; needs the diagnostic build: make -C icp_variants_b200/csrc groupprobe and
ICP_GPU_LIB_NAME=libicp_gpu_groupprobe.so) and the registration time with and without it, for pairs of different far-query shares."""
import ctypes as C, json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi

out = {}
gen = capi.Context(0)
for k in [int(x) for x in os.environ.get("PAIRS", "0,34").split(",")]:
    src, tgt = bench.make_pair_device_normals(gen, k, 344, 1077)
    r = {}
    for gm in [int(x) for x in os.environ.get("GROUP_MINS", "0,8").replace(",", " ").split()]:
        os.environ["ICP_GPU_GROUP_MIN"] = str(gm)
        ctx = capi.Context(0)                              # a fresh context: the knob is read when the loop is captured
        cfg = capi.default_config(); cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, 30, 10.0, 2, 0
        ctx.set_config(cfg); ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
        best = None
        for _ in range(4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pose, _, _ = ctx.estimate_pose(want_history=False); e1.record(); e1.synchronize()
            best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
        for _ in range(2):
            _, _, _, tm = ctx.estimate_pose(want_history=False, timings=True)
        cfg.collect_stats = 1; ctx.set_config(cfg)
        ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
        g = (C.c_ulonglong * 12)()
        probe = hasattr(capi.lib(), "icp_gpu_debug_group_stats")
        if probe:
            capi.lib().icp_gpu_debug_group_stats(g, 1)
        ctx.estimate_pose(want_history=False); st = ctx.stats()
        if probe:
            capi.lib().icp_gpu_debug_group_stats(g, 0)
        r[gm] = {"ms_30_iterations": best, "prep_us": tm.search_prep_ms / 30 * 1e3, "group_and_walk_us": (tm.matching_ms - tm.search_prep_ms) / 30 * 1e3,
                 "pose_checksum": float(np.abs(pose).sum()), "evals_per_launch": st.n_distance_evals / 30, "nodes_per_launch": st.n_nodes_visited / 30,
                 "per_iteration": {"windows": g[0] / 30, "groups_started": g[1] / 30, "groups_finished": g[2] / 30, "members_finished": g[3] / 30,
                                   "leaves_listed_per_group": g[4] / max(g[2], 1), "leaf_scans_per_group": g[5] / max(g[2], 1),
                                   "cycles_bounds_per_group": g[6] / max(g[2], 1), "cycles_descent_per_group": g[7] / max(g[2], 1),
                                   "cycles_scans_per_group": g[8] / max(g[2], 1), "cycles_longest_group": g[9], "box_rounds_per_group": g[10] / max(g[2], 1)}}
        ctx.close()
    out[k] = r
print(json.dumps(out, indent=1))
