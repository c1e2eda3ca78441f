"""The 44-pair queue (bench.py: pair_queue_44) replayed on ONE GPU: the whole queue and every rank's share of the 8-rank deal
(parallel.shard_pairs), with 3 / 4 / 6 / 8 contexts per GPU, page-locked host arrays in, poses out; and every pair alone.
Says how much of the 8-GPU figure is the short queue (ramp + tail of 5-6 pairs on three contexts) and how much the pairs' costs."""
import json, os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi, parallel, sequence, synth

N_PAIRS = int(os.environ.get("PAIRS", bench.N_SEQUENCE_PAIRS))
gen = capi.Context(0)
cfg = capi.default_config(); cfg.metric, cfg.n_iterations, cfg.max_distance_sq, cfg.nn_algorithm, cfg.collect_stats = 1, 30, 10.0, 2, 0
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()          # noqa: E731
pairs = []
for k in range(N_PAIRS):
    pr = bench.make_pair_device_normals(gen, k, 344, 1077)
    pairs.append(tuple(synth.Cloud(pin(c.points), pin(c.normals), pin(c.colors)) for c in pr[:2]))


def run(ctxs, share, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = sequence.alignPairs(ctxs, share, cfg)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        assert all(r.error is None and r.nIterations == 30 for r in res)
        best = dt if best is None else min(best, dt)
    return best


out = {"pairs": N_PAIRS, "whole_queue_ms": {}, "deal_of_8_ranks_ms": {}, "deal_of_4_ranks_ms": {}}
one = [capi.Context(0)]
out["ms_per_pair_alone"] = [round(run(one, [pairs[k]], reps=2), 3) for k in range(N_PAIRS)]
one[0].close()
for n_ctx in (3, 4, 6, 8):
    ctxs = [capi.Context(0) for _ in range(n_ctx)]
    out["whole_queue_ms"][n_ctx] = round(run(ctxs, pairs), 2)
    for world, key in ((8, "deal_of_8_ranks_ms"), (4, "deal_of_4_ranks_ms")):
        out[key][n_ctx] = [round(run(ctxs, [pairs[k] for k in parallel.shard_pairs(N_PAIRS, world, r)]), 2) for r in range(world)]
    for c in ctxs:
        c.close()
print(json.dumps(out, indent=1))
