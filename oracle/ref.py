"""ctypes front-end of oracle/_ref/libicp_ref.so: the reference's OWN headers
(/root/reference/icp-variants/*.h) compiled in place against the stand-ins in oracle/ref_shim/.

TEST INFRASTRUCTURE ONLY -- used by tests/ (to pin oracle/icp_oracle.c against the reference's own
code) and by tests/golden/make_reference_fixtures.py.  The library is built where /root/reference
exists (this container); on the GPU box the prebuilt .so travels with the snapshot.  What it pins
and what it does not is stated in oracle/ref_driver.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libicp_ref.so")
REFERENCE_ROOT = os.environ.get("ICP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE_ROOT, "icp-variants"))


def build(force: bool = False) -> str | None:
    """Compile the reference where its sources exist; otherwise use the prebuilt library (or None)."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "icp-variants")):
        subprocess.run(["make", "-C", _HERE, "ref", f"REFERENCE_ROOT={REFERENCE_ROOT}"] + (["-B"] if force else []),
                       check=True, stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


class _Config(C.Structure):
    _fields_ = [("minimizer", C.c_int32), ("metric", C.c_int32), ("selection", C.c_int32), ("weighting", C.c_int32),
                ("rejection", C.c_int32), ("matching", C.c_int32), ("color_icp", C.c_int32), ("multires", C.c_int32),
                ("n_iterations", C.c_int32), ("proba", C.c_double), ("seed", C.c_uint32), ("max_distance_sq", C.c_float),
                ("K", C.c_float * 9), ("width", C.c_uint32), ("height", C.c_uint32), ("setter_order", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libicp_ref.so is missing and /root/reference is not present")
        _lib = C.CDLL(so)
        for f in ("ref_coarse_resolution", "ref_selection", "ref_cloud_from_off", "ref_cloud_from_depth", "ref_cloud_from_xyz"):
            getattr(_lib, f).restype = C.c_int64
        _lib.ref_rmse.restype = C.c_float
        _lib.ref_benchmark_error.restype = C.c_double
        _lib.ref_describe.restype = C.c_char_p
    return _lib


def _f32(a, cols=3):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pose(p):
    return np.ascontiguousarray(np.asarray(p, dtype=np.float32).T.reshape(16))


def _unpose(v):
    return np.asarray(v, dtype=np.float32).reshape(4, 4).T.copy()


def _K9(fx, fy, cx, cy):
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    return np.ascontiguousarray(K.T.reshape(9))


def transform_points(pose, pts):
    pts = _f32(pts); out = np.empty_like(pts)
    lib().ref_transform_points(_p(_pose(pose)), _p(pts), C.c_int64(len(pts)), _p(out))
    return out


def transform_normals(pose, nrm):
    nrm = _f32(nrm); out = np.empty_like(nrm)
    lib().ref_transform_normals(_p(_pose(pose)), _p(nrm), C.c_int64(len(nrm)), _p(out))
    return out


def knn_flann(tgt, qry, max_d2, tgt_rgba=None, qry_rgba=None):
    tgt, qry = _f32(tgt), _f32(qry); tc, qc = _u8(tgt_rgba), _u8(qry_rgba)
    idx = np.empty(len(qry), np.int32); w = np.empty(len(qry), np.float32)
    rc = lib().ref_knn_flann(_p(tgt), _p(tc), C.c_int64(len(tgt)), _p(qry), _p(qc), C.c_int64(len(qry)), C.c_float(max_d2), _p(idx), _p(w))
    assert rc == 0
    return idx, w


def knn_brute(tgt, qry, max_d):
    tgt, qry = _f32(tgt), _f32(qry)
    idx = np.empty(len(qry), np.int32); w = np.empty(len(qry), np.float32)
    rc = lib().ref_knn_brute(_p(tgt), C.c_int64(len(tgt)), _p(qry), C.c_int64(len(qry)), C.c_float(max_d), _p(idx), _p(w))
    assert rc == 0
    return idx, w


def projective(tgt, width, height, fx, fy, cx, cy, qry, max_d2):
    tgt, qry = _f32(tgt), _f32(qry)
    idx = np.empty(len(qry), np.int32); w = np.empty(len(qry), np.float32)
    rc = lib().ref_projective(_p(tgt), C.c_uint32(width), C.c_uint32(height), _p(_K9(fx, fy, cx, cy)), _p(qry), C.c_int64(len(qry)),
                              C.c_float(max_d2), _p(idx), _p(w))
    assert rc == 0
    return idx, w


def apply_weights(method, max_d2, sp, sn, sc, tp, tn, tc, idx, w):
    sp, sn, tp, tn = _f32(sp), _f32(sn), _f32(tp), _f32(tn); sc, tc = _u8(sc), _u8(tc)
    idx = np.ascontiguousarray(idx, np.int32).copy(); w = np.ascontiguousarray(w, np.float32).copy()
    lib().ref_apply_weights(C.c_int(method), C.c_float(max_d2), _p(sp), _p(sn), _p(sc), C.c_int64(len(sp)),
                            _p(tp), _p(tn), _p(tc), C.c_int64(len(tp)), _p(idx), _p(w))
    return idx, w


def prune(sn, tn, idx, w):
    sn, tn = _f32(sn), _f32(tn)
    idx = np.ascontiguousarray(idx, np.int32).copy(); w = np.ascontiguousarray(w, np.float32).copy()
    lib().ref_prune(_p(sn), C.c_int64(len(sn)), _p(tn), C.c_int64(len(tn)), _p(idx), _p(w))
    return idx, w


def solve_linear(metric, s, d, ns, nt, w):
    s, d, ns, nt = _f32(s), _f32(d), _f32(ns), _f32(nt); w = np.ascontiguousarray(w, np.float32); out = np.empty(16, np.float32)
    rc = lib().ref_solve_linear(C.c_int(metric), _p(s), _p(d), _p(ns), _p(nt), _p(w), C.c_int64(len(s)), _p(out))
    return rc, _unpose(out)


def residuals(kind, x, s, d, ns, nt, w):
    x = np.ascontiguousarray(x, np.float64); out = np.zeros(3, np.float64)
    a = [np.ascontiguousarray(v, np.float32) for v in (s, d, ns, nt)]
    n = lib().ref_residuals(C.c_int(kind), _p(x), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), C.c_float(w), _p(out))
    return out[:n].copy()


def increment_to_matrix(x6):
    x = np.ascontiguousarray(x6, np.float64); out = np.empty(16, np.float32)
    lib().ref_increment_to_matrix(_p(x), _p(out))
    return _unpose(out)


def coarse_resolution(pts, nrm, rgba, factor):
    pts, nrm = _f32(pts), _f32(nrm); rgba = _u8(rgba)
    po = np.empty_like(pts); no = np.empty_like(nrm); co = np.empty((len(pts), 4), np.uint8)
    n = lib().ref_coarse_resolution(_p(pts), _p(nrm), _p(rgba), C.c_int64(len(pts)), C.c_int(factor), _p(po), _p(no), _p(co))
    return po[:n].copy(), no[:n].copy(), co[:n].copy()


def selection(pts, nrm, rgba, proba, seed, n_resamples=1):
    pts, nrm = _f32(pts), _f32(nrm); rgba = _u8(rgba)
    po = np.empty_like(pts); no = np.empty_like(nrm); nc = C.c_int64(0)
    n = lib().ref_selection(_p(pts), _p(nrm), _p(rgba), C.c_int64(len(pts)), C.c_double(proba), C.c_uint32(seed), C.c_int(n_resamples),
                            _p(po), _p(no), C.byref(nc))
    return po[:n].copy(), no[:n].copy(), nc.value


def rmse(pose, src, ref):
    src, ref = _f32(src), _f32(ref)
    return float(lib().ref_rmse(_p(_pose(pose)), _p(src), _p(ref), C.c_int64(len(src))))


def benchmark_error(pose, src, ref):
    src, ref = _f32(src), _f32(ref)
    return float(lib().ref_benchmark_error(_p(_pose(pose)), _p(src), _p(ref), C.c_int64(len(src))))


def cloud_from_off(path, cap=1 << 20):
    po = np.empty((cap, 3), np.float32); no = np.empty((cap, 3), np.float32)
    n = lib().ref_cloud_from_off(path.encode(), C.c_int64(cap), _p(po), _p(no))
    if n < 0:
        raise RuntimeError(f"ref_cloud_from_off({path}) -> {n}")
    return po[:n].copy(), no[:n].copy()


def cloud_from_depth(depth, rgba, fx, fy, cx, cy, extrinsics=None, keep_original_size=False, downsample=1, max_distance=0.1):
    depth = np.ascontiguousarray(depth, np.float32); h, w = depth.shape
    rgba = np.ascontiguousarray(rgba, np.uint8).reshape(-1)
    # PointCloud.h:151-152 reads colorFrame[i .. i+3] with the PIXEL index i (not 4*i); pad so the read stays in bounds
    assert rgba.size >= h * w + 3
    E = _pose(np.eye(4, dtype=np.float32) if extrinsics is None else extrinsics)
    po = np.empty((h * w, 3), np.float32); no = np.empty((h * w, 3), np.float32); co = np.empty((h * w, 4), np.uint8)
    n = lib().ref_cloud_from_depth(_p(depth), _p(rgba), _p(_K9(fx, fy, cx, cy)), _p(E), C.c_uint32(w), C.c_uint32(h),
                                   C.c_int(int(keep_original_size)), C.c_uint32(downsample), C.c_float(max_distance), _p(po), _p(no), _p(co))
    return po[:n].copy(), no[:n].copy(), co[:n].copy()


def cloud_from_xyz(pts):
    pts = _f32(pts); no = np.empty_like(pts); co = np.empty((len(pts), 4), np.uint8)
    n = lib().ref_cloud_from_xyz(_p(pts), C.c_int64(len(pts)), _p(no), _p(co))
    return no[:n].copy(), co[:n].copy()


def estimate_pose(minimizer, metric, src, src_n, src_c, tgt, tgt_n, tgt_c, gt_src, gt_ref, *, n_iterations=20, max_distance_sq=0.0003,
                  selection=0, proba=1.0, seed=0, weighting=0, rejection=1, matching=0, color_icp=False, multires=False,
                  camera=None, init_pose=None, setter_order=0):
    """Runs the reference's {Linear,Ceres}ICPOptimizer::estimatePose.  Returns (n_iterations_executed | -2, pose, rmse history).
    setter_order: 0 = setMatchingMethod then setMatchingMaxDistance (main.cpp); 1 = the other way round (the matcher falls back to
    MAX_DISTANCE, only the weighting keeps max_distance_sq); 2 = setMatchingMaxDistance never called."""
    src, src_n, tgt, tgt_n = _f32(src), _f32(src_n), _f32(tgt), _f32(tgt_n)
    src_c, tgt_c = _u8(src_c), _u8(tgt_c)
    gt_src, gt_ref = _f32(gt_src), _f32(gt_ref)
    cfg = _Config(minimizer, metric, selection, weighting, rejection, matching, int(color_icp), int(multires), n_iterations,
                  float(proba), seed & 0xFFFFFFFF, float(max_distance_sq))
    if camera is not None:
        fx, fy, cx, cy, width, height = camera
        for i, v in enumerate(_K9(fx, fy, cx, cy)):
            cfg.K[i] = float(v)
        cfg.width, cfg.height = width, height
    cfg.setter_order = int(setter_order)
    pose = _pose(np.eye(4, dtype=np.float32) if init_pose is None else init_pose).copy()
    cap = n_iterations + 128
    hist = np.zeros(cap, np.float32)
    n = lib().ref_estimate_pose(C.byref(cfg), _p(src), _p(src_n), _p(src_c), C.c_int64(len(src)), _p(tgt), _p(tgt_n), _p(tgt_c), C.c_int64(len(tgt)),
                                _p(gt_src), _p(gt_ref), C.c_int64(len(gt_src)), _p(pose), _p(hist), C.c_int(cap), None)
    return n, _unpose(pose), hist[:max(n, 0)].copy()


def set_num_threads(n: int):
    """Threads of the stand-in matcher's OpenMP loop (everything else in the reference's loop is single-threaded)."""
    lib().ref_set_num_threads(C.c_int(int(n)))


def num_threads() -> int:
    return int(lib().ref_num_threads())


def set_flann_exhaustive(on: bool):
    """Answer the FLANN stand-in's searches by the literal O(N*M) scan instead of its exact kd-tree (same results)."""
    lib().ref_set_flann_exhaustive(C.c_int(int(on)))


def describe():
    return lib().ref_describe().decode()
