"""Generates tests/golden/reference_outputs.npz from the reference's OWN code
(oracle/_ref/libicp_ref.so = /root/reference/icp-variants/*.h compiled in place against the
oracle/ref_shim stand-ins; see oracle/ref_driver.cpp).  Run in the build container, where
/root/reference exists:

    python tests/golden/make_reference_fixtures.py

The fixture lets the parity tests check the oracle AND the CUDA path against reference outputs
on machines where neither /root/reference nor the compiled library is available.  Inputs are
either the bundled bunny (tests/golden/bunny.npz) or seeded generators, so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from icp_variants_b200 import synth  # noqa: E402
from oracle import ref as R  # noqa: E402

VARIANTS = [("base", {}), ("random", dict(selection=1, proba=0.5, seed=7)), ("distance_weights", dict(weighting=1)),
            ("multires", dict(multires=True))]


def stage_inputs():
    """Seeded stage-level inputs shared by the generator and the tests."""
    rng = np.random.default_rng(20240601)
    tp = (np.round(rng.uniform(-1, 1, (2500, 3)) * 32) / 32).astype(np.float32)
    sp = (np.round(rng.uniform(-1, 1, (1500, 3)) * 32) / 32 + rng.normal(0, 0.01, (1500, 3))).astype(np.float32)
    sp[::97] = tp[:len(sp[::97])]                       # exact hits
    sp[5, 0] = np.nan; sp[11, 2] = -np.inf
    tn = rng.normal(size=(2500, 3)).astype(np.float32); tn /= np.linalg.norm(tn, axis=1, keepdims=True)
    sn = rng.normal(size=(1500, 3)).astype(np.float32); sn /= np.linalg.norm(sn, axis=1, keepdims=True)
    sn[7, 1] = -np.inf; tn[3, 0] = np.nan
    tc = rng.integers(0, 4, (2500, 4), dtype=np.uint8) * 64
    sc = rng.integers(0, 4, (1500, 4), dtype=np.uint8) * 64
    pose = synth.make_pose([0.02, -0.01, 0.03], [2.0, -1.0, 3.0])
    return sp, sn, sc, tp, tn, tc, pose


def main():
    out = {}
    src, tgt, gs, gt = synth.load_bunny()
    for minimizer in (0, 1):
        for metric in (0, 1, 2):
            for name, kw in VARIANTS:
                n, pose, rmse = R.estimate_pose(minimizer, metric, src.points, src.normals, src.colors, tgt.points, tgt.normals,
                                                tgt.colors, src.points[gs], tgt.points[gt], n_iterations=20, max_distance_sq=0.0003, **kw)
                assert n == 20
                out[f"bunny_{minimizer}_{metric}_{name}_pose"] = pose
                out[f"bunny_{minimizer}_{metric}_{name}_rmse"] = rmse
    sp, sn, sc, tp, tn, tc, pose = stage_inputs()
    q = R.transform_points(pose, sp)
    qn = R.transform_normals(pose, sn)
    out["stage_tp"] = q; out["stage_tn"] = qn
    for max_d2 in (0.002, 10.0):
        i3, w3 = R.knn_flann(tp, q, max_d2)
        i6, w6 = R.knn_flann(tp, q, max_d2, tc, sc)
        out[f"stage_knn3_{max_d2}"] = i3; out[f"stage_knn6_{max_d2}"] = i6
        for method in (1, 2, 3):
            i, w = R.apply_weights(method, max_d2, q, qn, sc, tp, tn, tc, i3, w3)
            i, w = R.prune(qn, tn, i, w)
            out[f"stage_idx_{method}_{max_d2}"] = i; out[f"stage_w_{method}_{max_d2}"] = w
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays;", R.describe())


if __name__ == "__main__":
    main()
