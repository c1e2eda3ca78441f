"""Stage split of BASELINE config 4 (coloured ETH-shaped pair, multires + symmetric + LM + 6-D k-NN + colour weighting)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import capi, synth
srcc, tgtc, _ = synth.eth_pair(seed=1234, colors="texture")
ctx = capi.Context(0)
out = {}
for name, minimizer, multires in (("LM_multires", 1, 1), ("LM_full", 1, 0), ("linear_full", 0, 0)):
    cfg = capi.default_config(); cfg.collect_stats = 0
    cfg.metric, cfg.minimizer, cfg.weighting, cfg.color_icp, cfg.multires, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm = 2, minimizer, 3, 1, multires, 30, 0.1, 2
    ctx.set_config(cfg); ctx.set_target(tgtc.points, tgtc.normals, tgtc.colors); ctx.set_source(srcc.points, srcc.normals, srcc.colors)
    ctx.estimate_pose()
    pose, hist, n_it, tm = ctx.estimate_pose(timings=True)
    out[name] = {"iterations": n_it, "matching_ms": tm.matching_ms, "prep_ms": tm.search_prep_ms, "solver_ms": tm.solver_ms, "total_ms": tm.total_ms,
                 "match_launches": tm.n_match_launches, "solver_launches": tm.n_solver_launches}
print(json.dumps(out, indent=1))
