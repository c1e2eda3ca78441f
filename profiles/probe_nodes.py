import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi
src,tgt=bench.make_pair(0)
ctx=capi.Context(0)
cfg=capi.default_config(); cfg.metric=1; cfg.max_distance_sq=10.0; cfg.nn_algorithm=2
ctx.set_config(cfg); ctx.set_target(tgt.points,tgt.normals,tgt.colors); ctx.set_source(src.points,src.normals,src.colors)
I=np.eye(4,dtype=np.float32)
for k in range(3):
    idx,w=ctx.query_matches(I); st=ctx.stats()
    print('call',k,'nodes/query %.2f evals/query %.1f'%(st.n_nodes_visited/len(src), st.n_distance_evals/len(src)))
