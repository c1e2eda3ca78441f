"""Times one registration of each BASELINE.json config on cuda:0 (CUDA events around estimate_pose, clouds
resident, median of 5 after 2 warm-ups).  Not a bench line: context for DESIGN.md.  Usage: python profiles/measure_configs.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import capi, synth  # noqa: E402


def timed(ctx, n=5, warm=2, **kw):
    stream = torch.cuda.current_stream()
    ts = []
    for i in range(warm + n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pose, _, n_it = ctx.estimate_pose(want_history=False, **kw)
        e1.record(stream)
        e1.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), n_it


def main():
    torch.cuda.set_device(0)
    ctx = capi.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    out = {}
    # C1 bunny, p2p linear, 20 iterations
    src, tgt, _, _ = synth.load_bunny()
    cfg = capi.default_config(); cfg.collect_stats = 0
    ctx.set_config(cfg); ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
    out["C1_bunny_p2p_linear_20it"] = timed(ctx)
    # C2 with the experiment runner's max distance (0.1) next to the driver's (10)
    src, tgt, _ = synth.eth_pair(seed=1234)
    for md in (10.0, 0.1):
        cfg = capi.default_config(); cfg.collect_stats = 0
        cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm = 1, 30, md, 2
        ctx.set_config(cfg); ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
        out[f"C2_eth370k_p2plane_linear_30it_maxd2_{md}"] = timed(ctx)
    # C4 coloured ETH-shaped pair, multires + symmetric + LM + 6-D k-NN + colour weighting
    srcc, tgtc, _ = synth.eth_pair(seed=1234, colors="texture")
    cfg = capi.default_config(); cfg.collect_stats = 0
    cfg.metric, cfg.minimizer, cfg.weighting, cfg.color_icp, cfg.multires, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm = 2, 1, 3, 1, 1, 30, 0.1, 2
    ctx.set_config(cfg); ctx.set_target(tgtc.points, tgtc.normals, tgtc.colors); ctx.set_source(srcc.points, srcc.normals, srcc.colors)
    out["C4_eth370k_color6d_multires_symmetric_LM_30it"] = timed(ctx)
    # C3 TUM-shaped 640x480, projective + normals weighting + symmetric linear, 35 iterations
    s3, t3, k, _ = synth.tum_pair(seed=1234, frame_gap=10)
    cfg = capi.default_config(); cfg.collect_stats = 0
    cfg.metric, cfg.matching, cfg.weighting, cfg.n_iterations, cfg.max_distance_sq = 2, 1, 2, 35, 0.1
    ctx.set_config(cfg); ctx.set_camera(k, 640, 480)
    ctx.set_target(t3.points, t3.normals, t3.colors); ctx.set_source(s3.points, s3.normals, s3.colors)
    out["C3_tum640x480_projective_normalsw_symmetric_35it"] = timed(ctx)
    print(json.dumps({k: {"ms": v[0], "iterations": v[1]} for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
