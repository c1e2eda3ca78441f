import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import capi, synth
from oracle import oracle as orc
src, tgt, _ = synth.eth_pair(seed=1234, n_sweeps=60, n_beams=160)
ctx = capi.Context(0)
for nit in (1, 2, 3):
    cfg = capi.default_config(); cfg.metric = 1; cfg.multires = 1; cfg.pyramid_mode = 1; cfg.max_distance_sq = 0.1; cfg.n_iterations = nit; cfg.nn_algorithm = 2
    ctx.set_config(cfg); ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
    pose, hist, n_it = ctx.estimate_pose()
    print('n_iterations', nit, 'executed', n_it, 'device n_queries', ctx.stats().n_queries)
sizes = {s: len(orc.voxel_indices(src.points, src.normals, s)) for s in (64, 32, 16, 8, 4, 2, 1)}
print('oracle level sizes', sizes, 'sum', sum(sizes.values()))
