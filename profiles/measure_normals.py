"""k = 5 PCA normals of the 370k-point ETH-shaped cloud on the device (icp_gpu_target_normals), CUDA-event timed; and the
depth -> cloud kernel chain at 640x480.  Context for DESIGN.md.  Usage: python profiles/measure_normals.py"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import capi, synth

torch.cuda.set_device(0)
ctx = capi.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
_, tgt, _ = synth.eth_pair(seed=1234)
ctx.set_target(tgt.points, None, None)
lib = capi.lib()
out = {}
ts = []
for i in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = lib.icp_gpu_target_normals(ctx._h, 5, None, None, None); e1.record(); e1.synchronize()
    assert rc == 0
    ts.append(e0.elapsed_time(e1))
out["pca_normals_k5_370k_ms"] = float(np.median(ts[2:]))
frames, K, _ = synth.tum_sequence(n_frames=1, seed=1)
ts = []
for i in range(7):
    t0 = time.perf_counter(); n = ctx.cloud_from_depth(frames[0], None, K, None, True, 1, 0.1, role=0, download=False); ts.append((time.perf_counter() - t0) * 1e3)
out["depth_640x480_to_indexed_target_ms_host_clock"] = float(np.median(ts[2:]))
print(json.dumps(out, indent=1))
