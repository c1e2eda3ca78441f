"""Where a frame of the TUM-shaped sequence spends its host-clock time (profiles/measure_sequence.py gives the total): the calls of
sequence.reconstructRoom's loop body timed one by one, each followed by a device synchronisation.  Context for DESIGN.md.
Usage: python profiles/probe_sequence_frame.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icp_variants_b200 import synth  # noqa: E402
from icp_variants_b200.optimizer import ConvergenceMeasure, LinearICPOptimizer  # noqa: E402

frames, K, gt = synth.tum_sequence(n_frames=11, seed=1234)
frames = np.asarray(frames, np.float32)
out = {}
for name, projective in (("projective", True), ("knn", False)):
    opt = LinearICPOptimizer(device=0)
    opt.setMetric(2); opt.setNbOfIterations(35)
    if projective:
        opt.setMatchingMethod(1)
        opt.setCameraParamsMatchingMethod(K, frames.shape[2], frames.shape[1])
    opt.setMatchingMaxDistance(0.1)
    opt.setWeightingMethod(2 if projective else 0)
    opt.setTargetFromDepth(frames[0], None, K, None, keepOriginalSize=projective, maxDistance=0.1)
    rows = []
    cur = np.eye(4, dtype=np.float32)
    for rep in range(2):
        for i in range(1, len(frames)):
            t0 = time.perf_counter()
            opt.setSourceFromDepth(frames[i], None, K, None, keepOriginalSize=False, downsampleFactor=8, maxDistance=0.1)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            cm = ConvergenceMeasure(groundTruthPose=np.asarray(gt[i], np.float32))
            opt.setConvergenceMeasure(cm)
            opt._ctx.set_config(opt.config())
            if opt._camera is not None:
                opt._ctx.set_camera(*opt._camera)
            t3 = time.perf_counter()
            res = opt._ctx.estimate_pose(cur, want_history=True, timings=False)
            t4 = time.perf_counter()
            opt._ctx.set_correspondences_pose(cm.groundTruthPose)
            t5 = time.perf_counter()
            opt._ctx.convergence_errors(benchmark=False)
            t6 = time.perf_counter()
            if rep == 1:
                rows.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5])
    med = np.median(np.asarray(rows), axis=0) * 1e3
    out[name] = dict(zip(["set_source_from_depth_ms", "sync_after_it_ms", "set_config_ms", "estimate_pose_ms", "set_correspondences_pose_ms",
                          "convergence_errors_ms"], [round(float(x), 4) for x in med]))
print(json.dumps(out, indent=1))
