"""oracle/flann_like.py -- TEST / MEASUREMENT INFRASTRUCTURE (never imported by the product path).

The reference's matcher is FLANN 1.8.4's randomized kd-tree with ONE tree and 16 checks
(`flann::Index<flann::L2<float>>(…, flann::KDTreeIndexParams(1))`, `flann::SearchParams(16)`,
/root/reference/icp-variants/NearestNeighbor.h:136,172-174): an APPROXIMATE search.  FLANN itself is neither vendored
nor installed here, so the closest available stand-in is OpenCV's fork of the same library, `cv2.flann_Index` with
`algorithm = FLANN_INDEX_KDTREE (1)`, `trees = 1`, `checks = 16`.  This module runs the reference's
linear point-to-plane loop with that matcher -- every other stage is the oracle's restatement (transformPoints /
transformNormals utils.h:106-133, the d2 <= max threshold NearestNeighbor.h:182, pruneCorrespondences
ICPOptimizer.h:157-174, estimatePosePointToPlane :676-782) -- and reports

  * seconds per iteration on the host (index build counted once), and
  * the MATCH RATE: the fraction of queries whose FLANN-like answer equals the exact nearest neighbour (the oracle's
    kd-tree, == brute force with lowest-index ties) at the same pose -- north_star reports FLANN's approximate matches only
    as this rate; parity is defined against the exact search.

parity unpinned for this file's matcher: OpenCV's FLANN fork draws its own random split dimensions, so individual answers
differ from FLANN 1.8.4's; only the rate is meaningful.
"""
from __future__ import annotations

import time

import numpy as np

from . import oracle as orc

FLANN_INDEX_KDTREE = 1


def available() -> bool:
    try:
        import cv2  # noqa: F401
        return hasattr(cv2, "flann_Index")
    except Exception:   # noqa: BLE001
        return False


def register_p2plane(src, tgt, max_d2: float, iterations: int, init_pose=None, seed: int = 0):
    """Point-to-plane linear ICP (constant weights, normal-angle rejection on) with the FLANN-like matcher.
    Returns dict(pose, seconds_build, seconds_per_iteration [list], match_rate [list], matched_fraction [list])."""
    import cv2
    cv2.setRNGSeed(seed)
    tp = np.ascontiguousarray(tgt.points, np.float32)
    t0 = time.perf_counter()
    index = cv2.flann_Index(tp, {"algorithm": FLANN_INDEX_KDTREE, "trees": 1})       # NearestNeighbor.h:136
    t_build = time.perf_counter() - t0
    exact = orc.KdTree(tp)
    pose = np.eye(4, dtype=np.float32) if init_pose is None else np.asarray(init_pose, np.float32)
    secs, rate, frac = [], [], []
    n = len(src.points)
    for _ in range(iterations):
        t0 = time.perf_counter()
        q = orc.transform_points(pose, src.points)                                     # ICPOptimizer.h:553
        qn = orc.transform_normals(pose, src.normals)                                  # :554
        fin = np.isfinite(q).all(1)
        idx = np.full(n, -1, np.int32); d2 = np.full(n, np.inf, np.float32)
        if fin.any():
            i, d = index.knnSearch(np.ascontiguousarray(q[fin]), 1, params={"checks": 16})   # NearestNeighbor.h:172-174
            idx[fin] = i[:, 0]; d2[fin] = d[:, 0]
        m = np.zeros(n, orc.MATCH_DTYPE)
        ok = d2 <= np.float32(max_d2)                                                  # :182
        m["idx"] = np.where(ok, idx, -1); m["weight"] = ok.astype(np.float32)
        m = orc.prune(qn, tgt.normals, m)                                              # ICPOptimizer.h:578-579
        keep = m["idx"] >= 0
        rc, inc = orc.solve_p2plane(q[keep], tgt.points[m["idx"][keep]], tgt.normals[m["idx"][keep]], m["weight"][keep])   # :676-782
        secs.append(time.perf_counter() - t0)
        # the exact answer at the same pose (outside the timed part)
        e = exact.query(q, max_d2)["idx"]
        a = np.where(ok, idx, -1)
        rate.append(float(np.mean(a == e)))
        frac.append(float(np.mean(e >= 0)))
        if rc != 0:
            break
        pose = (inc.astype(np.float32) @ pose).astype(np.float32)                      # :614-620
    return {"pose": pose, "seconds_build": t_build, "seconds_per_iteration": secs, "match_rate": rate, "matched_fraction": frac}
