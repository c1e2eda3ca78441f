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


@dataclass
class PairResult:
    pose: np.ndarray
    nIterations: int
    rmseErrors: list = field(default_factory=list)
    benchmarkErrors: list = field(default_factory=list)
    error: str | None = None


def alignPairs(contexts, pairs, config, calculateErrors: bool = False, tickets=None):
    """The pair loop of alignETH (main.cpp:343-514; experiment.cpp:276-412): independent (source, target[, unchanged source])
    registrations, identity start pose, dealt round-robin to the given contexts (one or two per GPU, any number of GPUs).
    Registration k+1 is uploaded and enqueued on the next context before registration k is waited for
    (icp_gpu_estimate_pose_async / _finish), so uploads, index builds and loops of different contexts overlap.
    pairs: sequence of (source Cloud, target Cloud) or (source, target, unchangedSourcePoints [N,3]) -- the third entry feeds
    ConvergenceMeasure(source points, unchanged points, runBenchmark=true) as main.cpp:439 does.  Returns [PairResult].
    tickets (parallel.PairTickets or anything with next() -> index | None): several processes share ONE queue -- `pairs` is
    the whole sequence on every process, the index of the next pair is drawn when a context is free, and the entries of the
    result that other processes took stay None."""
    from . import capi
    pairs = list(pairs)
    results: list = [None] * len(pairs)
    pending: list = []                      # (pair index, context)

    def finish(k, ctx):
        try:
            pose, n_it = ctx.estimate_pose_finish()
            r = PairResult(pose, n_it)
        except capi.IcpGpuError as e:
            if e.code not in (capi.E_NO_MATCHES, capi.E_NUMERIC):
                raise
            # the reference spins in ASSERT / returns a NaN pose here; the drop-in reports it and goes on with the next pair
            r = PairResult(getattr(e, "pose", np.eye(4, dtype=np.float32)), 0, error=str(e))
        if calculateErrors and len(pairs[k]) > 2 and r.nIterations > 0:
            ctx.set_correspondences(pairs[k][0].points, pairs[k][2])
            rm, be = ctx.convergence_errors(benchmark=True)
            r.rmseErrors, r.benchmarkErrors = [float(x) for x in rm], [float(x) for x in be]
        results[k] = r

    n_enqueued, k_static = 0, 0
    while True:
        ctx = contexts[n_enqueued % len(contexts)]
        # a context still busy with an earlier pair must be drained first (and only then is the next ticket drawn: a process
        # that waits holds no pair back from the others)
        for j, (kk, cc) in enumerate(pending):
            if cc is ctx:
                finish(kk, cc)
                pending.pop(j)
                break
        if tickets is not None:
            k = tickets.next()
        else:
            k, k_static = (k_static, k_static + 1) if k_static < len(pairs) else (None, k_static)
        if k is None:
            break
        src, tgt = pairs[k][0], pairs[k][1]
        ctx.set_config(config)
        ctx.set_target(tgt.points, tgt.normals, tgt.colors)
        ctx.set_source(src.points, src.normals, src.colors)
        ctx.estimate_pose_async(np.eye(4, dtype=np.float32))
        pending.append((k, ctx))
        n_enqueued += 1
    for kk, cc in pending:
        finish(kk, cc)
    return results
