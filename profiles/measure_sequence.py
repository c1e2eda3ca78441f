"""Frames per second of the TUM-shaped sequence (BASELINE config 3 as reconstructRoom runs it, main.cpp:183-341): 640x480 frames,
projective matching + normals weighting + symmetric linear ICP, 35 iterations, pose carried over, RMSE per iteration against the
ground-truth trajectory -- every per-frame step on the device (icp_variants_b200/sequence.py).  Context for DESIGN.md, not a bench line.
Usage: python profiles/measure_sequence.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import synth  # noqa: E402
from icp_variants_b200.optimizer import LinearICPOptimizer  # noqa: E402
from icp_variants_b200.sequence import reconstructRoom  # noqa: E402

frames, K, gt = synth.tum_sequence(n_frames=11, seed=1234)
out = {}
for name, projective, multires in (("projective_normalsw_symmetric_35it", True, False), ("knn_stride8_symmetric_35it", False, False)):
    opt = LinearICPOptimizer(device=0)
    opt.setMetric(2); opt.setNbOfIterations(35)
    if projective:
        opt.setMatchingMethod(1)
    opt.setMatchingMaxDistance(0.1)
    opt.setWeightingMethod(2 if projective else 0)
    reconstructRoom(opt, frames[:3], K, groundTruthPoses=gt[:3])          # warm-up (graph capture, allocations)
    t0 = time.perf_counter()
    res = reconstructRoom(opt, frames, K, groundTruthPoses=gt)
    dt = time.perf_counter() - t0
    out[name] = {"frames": len(frames) - 1, "seconds": dt, "frames_per_s": (len(frames) - 1) / dt,
                 "ms_per_frame_median": float(np.median(res.secondsPerFrame) * 1e3), "source_points": res.nSourcePoints,
                 "final_rmse": [float(x) for x in res.finalRMSE]}
print(json.dumps(out, indent=1))
