"""Work per query of the BVH search, split by how far the query is from the target (ideal seeds: second call)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi
from scipy.spatial import cKDTree
src, tgt = bench.make_pair(0)
d, _ = cKDTree(tgt.points).query(src.points)
ctx = capi.Context(0)
cfg = capi.default_config(); cfg.metric = 1; cfg.max_distance_sq = 10.0; cfg.nn_algorithm = 2
ctx.set_config(cfg); ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
I = np.eye(4, dtype=np.float32)
for name, sel in (("all", None), ("d<5cm", np.where(d < 0.05)[0]), ("5-20cm", np.where((d >= 0.05) & (d < 0.2))[0]),
                  ("20cm-1m", np.where((d >= 0.2) & (d < 1.0))[0]), (">1m", np.where(d >= 1.0)[0])):
    for k in range(2):
        idx, w = ctx.query_matches(I, None if sel is None else sel.astype(np.int32)); st = ctx.stats()
        n = len(src) if sel is None else len(sel)
        print(f"{name:8s} call {k} queries {n:7d} nodes/query {st.n_nodes_visited / n:7.2f} evals/query {st.n_distance_evals / n:7.1f}")
