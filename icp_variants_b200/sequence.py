"""Frame-to-model tracking of an RGB-D sequence -- the caller one level above the registration loop
(reconstructRoom, main.cpp:183-341; experiment.cpp:143-274), with every per-frame step on the device:

    frame 0            -> PointCloud(depth, ...) built and indexed on the device        (main.cpp:202-206)
    frame i            -> PointCloud(depth, ...) built on the device as the source      (main.cpp:291-298)
    ConvergenceMeasure -> every source point under the ground-truth pose               (main.cpp:300-308)
    estimatePose       -> the loop, started from the previous frame's result            (main.cpp:312; pose carry-over)
    RMSE per iteration -> evaluated on the device from the pose history                 (main.cpp:315-322)

Per frame the host sends one depth map (+ optionally the RGBX frame) and a pose, and receives a pose and the errors.
Nothing here falls back to the CPU."""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from .optimizer import ConvergenceMeasure, ICPOptimizer


@dataclass
class SequenceResult:
    estimatedPoses: list = field(default_factory=list)      # camera poses, estimatedPoses.push_back(currentCameraToWorld.inverse())
    cameraToWorld: list = field(default_factory=list)       # the registration result per frame (source -> frame 0)
    finalRMSE: list = field(default_factory=list)
    rmsePerIteration: list = field(default_factory=list)
    nSourcePoints: list = field(default_factory=list)
    secondsPerFrame: list = field(default_factory=list)


def reconstructRoom(optimizer: ICPOptimizer, depthFrames, depthIntrinsics, colorFrames=None, groundTruthPoses=None, depthExtrinsics=None,
                    projective: bool | None = None, multiResolution: bool | None = None, maxDistance: float = 0.1) -> SequenceResult:
    """Tracks depthFrames[1:] against depthFrames[0] (the fixed target).  The optimizer carries the options
    (metric, iterations, matching / weighting / selection ...) exactly as main.cpp:210-268 sets them.
    groundTruthPoses[i] (optional) maps frame i's camera space to frame 0's (targetTrajectory * trajectory_i^-1)."""
    frames = np.asarray(depthFrames, np.float32)
    n, h, w = frames.shape
    projective = optimizer.matchingMethod == 1 if projective is None else projective
    multires = optimizer.multiResolutionICP if multiResolution is None else multiResolution
    if projective:
        optimizer.setCameraParamsMatchingMethod(depthIntrinsics, w, h)                      # main.cpp:236-239
    col = (lambda i: None) if colorFrames is None else (lambda i: colorFrames[i])
    # For projective search keep the whole target point cloud, even the invalid points (main.cpp:197-202)
    optimizer.setTargetFromDepth(frames[0], col(0), depthIntrinsics, depthExtrinsics, keepOriginalSize=projective, maxDistance=maxDistance)
    res = SequenceResult()
    current = np.eye(4, dtype=np.float32)                                                  # currentCameraToWorld
    res.estimatedPoses.append(np.linalg.inv(current).astype(np.float32))
    for i in range(1, n):
        t0 = time.perf_counter()
        # For multiresolution keep all the points, else every 8th valid one (main.cpp:291-298)
        ns = optimizer.setSourceFromDepth(frames[i], col(i), depthIntrinsics, depthExtrinsics, keepOriginalSize=multires,
                                          downsampleFactor=1 if multires else 8, maxDistance=maxDistance)
        cm = None
        if groundTruthPoses is not None:
            cm = ConvergenceMeasure(groundTruthPose=np.asarray(groundTruthPoses[i], np.float32))
            optimizer.setConvergenceMeasure(cm)
        pose = optimizer.estimatePose(None, None, current, calculateRMSE=cm is not None)
        if cm is not None:
            res.rmsePerIteration.append(list(cm.rmseErrors))
            res.finalRMSE.append(cm.rmseErrors[-1] if cm.rmseErrors else float("nan"))
        current = np.asarray(pose, np.float32)
        res.cameraToWorld.append(current.copy())
        res.estimatedPoses.append(np.linalg.inv(current.astype(np.float64)).astype(np.float32))
        res.nSourcePoints.append(int(ns))
        res.secondsPerFrame.append(time.perf_counter() - t0)
    return res
