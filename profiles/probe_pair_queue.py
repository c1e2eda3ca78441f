"""How many contexts (each with its own stream) per GPU serve the pair queue best: sequence.alignPairs over 12 ETH-shaped
pairs (host arrays in, poses out), wall clock between device synchronisations."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi, sequence
dev = torch.device('cuda', 0)
gen = capi.Context(0)
cfg = capi.default_config(); cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, 30, 10.0, 2, 0
idx = list(range(12))
pairs = [bench.make_pair_device_normals(gen, k, 344, 1077) for k in idx]
for n_ctx in (1, 2, 3, 4, 6):
    ctxs = [capi.Context(0) for _ in range(n_ctx)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = sequence.alignPairs(ctxs, pairs, cfg)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(n_ctx, "contexts:", f"{dt*1e3/len(pairs):.2f} ms/pair")
    for c in ctxs: c.close()
