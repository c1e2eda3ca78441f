"""Why the pairs of the 44-pair sequence differ in cost (2.7 ... 13 ms): per pair the registration time (index built, clouds
resident), the stage times per iteration, the work counters, and -- at the converged pose, seeds warm -- node visits and distance
evaluations per query by the query's distance from the target.  PAIRS=0,28,34 selects the pairs."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi
from scipy.spatial import cKDTree

ctx = capi.Context(0)
out = {}
for k in [int(x) for x in os.environ.get("PAIRS", "0,28,34,40").split(",")]:
    src, tgt = bench.make_pair_device_normals(ctx, k, 344, 1077)
    cfg = capi.default_config(); cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, 30, 10.0, 2, 0
    ctx.set_config(cfg); ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pose, _, _ = ctx.estimate_pose(want_history=False); e1.record(); e1.synchronize()
        best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
    for _ in range(2):
        _, _, _, tm = ctx.estimate_pose(want_history=False, timings=True)
    cfg.collect_stats = 1; ctx.set_config(cfg)
    ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
    ctx.estimate_pose(want_history=False); st = ctx.stats()
    r = {"ms_30_iterations": best, "prep_us": tm.search_prep_ms / 30 * 1e3, "walk_us": (tm.matching_ms - tm.search_prep_ms) / 30 * 1e3,
         "reduce_us": tm.solver_ms / 30 * 1e3, "evals_per_launch": st.n_distance_evals / 30, "nodes_per_launch": st.n_nodes_visited / 30,
         "matched_per_launch": st.n_matched / 30, "by_distance": {}}
    moved = (src.points.astype(np.float64) @ pose[:3, :3].T.astype(np.float64)) + pose[:3, 3].astype(np.float64)
    d, nn = cKDTree(tgt.points.astype(np.float64)).query(moved)
    r["unique_neighbours_of_far_queries"] = int(len(np.unique(nn[d > 0.2])))
    for name, sel in (("all", None), ("d<5cm", d < 0.05), ("5-20cm", (d >= 0.05) & (d < 0.2)), ("20cm-1m", (d >= 0.2) & (d < 1.0)),
                      ("1m-3.16m", (d >= 1.0) & (d < 3.1622)), (">3.16m (no match)", d >= 3.1623)):
        idx = None if sel is None else np.where(sel)[0].astype(np.int32)
        n = len(src) if idx is None else len(idx)
        if n == 0:
            continue
        for _ in range(2):                                # the second call starts from the first one's neighbours
            ctx.query_matches(pose, idx); s2 = ctx.stats()
        r["by_distance"][name] = {"queries": n, "nodes_per_query": s2.n_nodes_visited / n, "evals_per_query": s2.n_distance_evals / n}
    out[k] = r
print(json.dumps(out, indent=1))
