"""Device time of the index build alone (buildIndex = icp_gpu_set_target_dev: pack, radix sort, tree levels, boxes, adjacency lists)
and of the source sort (icp_gpu_set_source_dev), CUDA-event timed on the context's stream, L2 flushed between repetitions.
Usage: python profiles/measure_build.py [sweeps beams]  ->  one JSON line."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icp_variants_b200 import capi

sweeps, beams = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (344, 1077)
src, tgt = bench.make_pair(0, sweeps, beams)
dev = torch.device("cuda", 0)
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
     (("sp", src.points), ("sn", src.normals), ("sc", src.colors), ("tp", tgt.points), ("tn", tgt.normals), ("tc", tgt.colors))}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timed(fn, reps=20, warm=3):
    out = []
    for r in range(reps + warm):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        ctx.synchronize()                 # joins the part of the build that runs on the context's second stream
        e1.record(stream)
        e1.synchronize()
        if r >= warm:
            out.append(e0.elapsed_time(e1))
    return float(np.median(out)), float(np.min(out))

t_med, t_min = timed(lambda: ctx.set_target_dev(d["tp"].data_ptr(), d["tn"].data_ptr(), d["tc"].data_ptr(), len(tgt)))
s_med, s_min = timed(lambda: ctx.set_source_dev(d["sp"].data_ptr(), d["sn"].data_ptr(), d["sc"].data_ptr(), len(src)))
def both():
    ctx.set_target_dev(d["tp"].data_ptr(), d["tn"].data_ptr(), d["tc"].data_ptr(), len(tgt))
    ctx.set_source_dev(d["sp"].data_ptr(), d["sn"].data_ptr(), d["sc"].data_ptr(), len(src))
b_med, b_min = timed(both)
n = len(tgt)
print(json.dumps({"n_target": n, "n_source": len(src), "target_index_build_ms": {"median": t_med, "min": t_min},
                  "source_sort_ms": {"median": s_med, "min": s_min}, "both_ms": {"median": b_med, "min": b_min},
                  "algorithmic_bytes_per_point": 140, "target_build_gbs": 140.0 * n / (t_med * 1e-3) / 1e9,
                  "note": "events on the context's stream around the call + join; includes the host's launch gaps (the call enqueues ~17 kernels)"}))
ctx.close()
