"""One C3 registration (TUM-shaped 640x480, projective + normals weighting + symmetric linear, 35 iterations) for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import capi, synth
s3, t3, k, _ = synth.tum_pair(seed=1234, frame_gap=10)
ctx = capi.Context(0)
cfg = capi.default_config(); cfg.collect_stats = int(os.environ.get("STATS", "0"))
cfg.metric, cfg.matching, cfg.weighting, cfg.n_iterations, cfg.max_distance_sq = 2, 1, 2, 35, 0.1
ctx.set_config(cfg); ctx.set_camera(k, 640, 480)
ctx.set_target(t3.points, t3.normals, t3.colors); ctx.set_source(s3.points, s3.normals, s3.colors)
for _ in range(2):
    pose, _, n = ctx.estimate_pose()
st = ctx.stats()
print(n, st.n_queries, st.n_distance_evals, st.n_matched, 'unstaged blocks (last run: 35 iterations x 2400 blocks):', st.n_nodes_visited)
if os.environ.get("TIMINGS"):
    # event-timed matching stage per iteration (timings mode) and the FP32 fraction it implies
    res = ctx.estimate_pose(timings=True)
    t = res[3]
    peak = ctx.measure_fp32_peak(1) if hasattr(ctx, "measure_fp32_peak") else capi.measure_fp32_peak(0, 1)
    ev = st.n_distance_evals / 35 if st.n_distance_evals else float("nan")      # the counters cover the last registration (35 launches)
    ms = t.matching_ms / t.n_iterations
    print({"projective_kernel_ms_per_launch_events": ms, "distance_evals_per_launch": ev, "tflops_non_fma": ev * 8 / (ms * 1e-3) / 1e12,
           "measured_fmul_fadd_peak_tflops": peak, "frac": ev * 8 / (ms * 1e-3) / 1e12 / peak})
