"""ctypes front-end of the CPU oracle (oracle/icp_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (icp_variants_b200/) never imports this.
Pinned against the reference's own code through oracle/ref.py (see oracle/icp_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libicp_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("icp_oracle.c", "icp_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


class _Config(C.Structure):
    _fields_ = [("metric", C.c_int32), ("minimizer", C.c_int32), ("matching", C.c_int32), ("selection", C.c_int32),
                ("proba", C.c_double), ("seed", C.c_uint32), ("weighting", C.c_int32), ("rejection", C.c_int32),
                ("max_distance_sq", C.c_float), ("color_icp", C.c_int32), ("multires", C.c_int32),
                ("n_iterations", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_uint32), ("height", C.c_uint32), ("nn_mode", C.c_int32), ("lm_max_iterations", C.c_int32),
                ("pyramid_mode", C.c_int32), ("weight_max_distance_sq", C.c_float)]


MATCH_DTYPE = np.dtype([("idx", np.int32), ("weight", np.float32)])

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_rmse.restype = C.c_float
        _lib.orc_kdtree_build.restype = C.c_void_p
        _lib.orc_mt_canonical.restype = C.c_double
        _lib.orc_coarse_indices.restype = C.c_int64
        _lib.orc_cloud_from_depth.restype = C.c_int64
        _lib.orc_benchmark_error.restype = C.c_double
    return _lib


@dataclass
class Config:
    """Mirrors the option setters of ICPOptimizer (ICPOptimizer.h:41-95), defaults of :29-31."""
    metric: int = 0
    minimizer: int = 0          # 0 LinearICPOptimizer, 1 CeresICPOptimizer
    matching: int = 0           # 0 k-NN, 1 projective
    selection: int = 0
    proba: float = 1.0
    seed: int = 0
    weighting: int = 0
    rejection: int = 1
    max_distance_sq: float = 0.0003
    color_icp: bool = False
    multires: bool = False
    n_iterations: int = 20
    fx: float = 0.0
    fy: float = 0.0
    cx: float = 0.0
    cy: float = 0.0
    width: int = 0
    height: int = 0
    nn_mode: int = 1            # 0 brute force, 1 exact kd-tree (same answers)
    lm_max_iterations: int = 10
    pyramid_mode: int = 0       # 0 stride pyramid (reference), 1 voxel levels (extension)
    weight_max_distance_sq: float = 0.0   # WeightingMethod's maxDistance when it differs from the matcher's (0 = the same)

    def c(self) -> _Config:
        return _Config(self.metric, self.minimizer, self.matching, self.selection, float(self.proba), self.seed & 0xFFFFFFFF,
                       self.weighting, self.rejection, float(self.max_distance_sq), int(self.color_icp), int(self.multires),
                       self.n_iterations, self.fx, self.fy, self.cx, self.cy, self.width, self.height, self.nn_mode,
                       self.lm_max_iterations, self.pyramid_mode, float(self.weight_max_distance_sq))


def _f32(a, cols=3):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


def _u8(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pose(p):
    """4x4 row-major numpy -> 16 floats column-major (Eigen Matrix4f storage)."""
    return np.ascontiguousarray(np.asarray(p, dtype=np.float32).T.reshape(16))


def _unpose(v):
    return np.asarray(v, dtype=np.float32).reshape(4, 4).T.copy()


def transform_points(pose, pts):
    pts = _f32(pts); out = np.empty_like(pts)
    lib().orc_transform_points(_p(_pose(pose)), _p(pts), C.c_int64(len(pts)), _p(out))
    return out


def transform_normals(pose, nrm):
    nrm = _f32(nrm); out = np.empty_like(nrm)
    lib().orc_transform_normals(_p(_pose(pose)), _p(nrm), C.c_int64(len(nrm)), _p(out))
    return out


def knn_brute(tgt, qry, max_d2, tgt_rgba=None, qry_rgba=None):
    tgt, qry = _f32(tgt), _f32(qry)
    out = np.empty(len(qry), MATCH_DTYPE)
    if tgt_rgba is None:
        lib().orc_knn3_brute(_p(tgt), C.c_int64(len(tgt)), _p(qry), C.c_int64(len(qry)), C.c_float(max_d2), _p(out))
    else:
        tc, qc = _u8(tgt_rgba), _u8(qry_rgba)
        lib().orc_knn6_brute(_p(tgt), _p(tc), C.c_int64(len(tgt)), _p(qry), _p(qc), C.c_int64(len(qry)), C.c_float(max_d2), _p(out))
    return out


def knn_brute_norm(tgt, qry, max_d):
    """NearestNeighborSearchBruteForce::getClosestPoint as written: rounded norms, a plain distance threshold."""
    tgt, qry = _f32(tgt), _f32(qry)
    out = np.empty(len(qry), MATCH_DTYPE)
    lib().orc_knn3_brute_norm(_p(tgt), C.c_int64(len(tgt)), _p(qry), C.c_int64(len(qry)), C.c_float(max_d), _p(out))
    return out


class KdTree:
    def __init__(self, tgt, tgt_rgba=None):
        self._tgt = _f32(tgt); self._tc = _u8(tgt_rgba)
        self._h = C.c_void_p(lib().orc_kdtree_build(_p(self._tgt), _p(self._tc), C.c_int64(len(self._tgt))))

    def query(self, qry, max_d2, qry_rgba=None):
        qry = _f32(qry); qc = _u8(qry_rgba)
        out = np.empty(len(qry), MATCH_DTYPE)
        lib().orc_kdtree_query(self._h, _p(qry), _p(qc), C.c_int64(len(qry)), C.c_float(max_d2), _p(out))
        return out

    def __del__(self):
        try:
            lib().orc_kdtree_free(self._h)
        except Exception:
            pass


def projective(tgt, width, height, fx, fy, cx, cy, qry, max_d2):
    tgt, qry = _f32(tgt), _f32(qry)
    out = np.empty(len(qry), MATCH_DTYPE)
    lib().orc_projective(_p(tgt), C.c_uint32(width), C.c_uint32(height), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                         _p(qry), C.c_int64(len(qry)), C.c_float(max_d2), _p(out))
    return out


def apply_weights(method, max_d2, sp, sn, sc, tp, tn, tc, matches):
    """weighting.h:39-99 on per-source arrays + matches (copy returned)."""
    sp, sn, tp, tn = _f32(sp), _f32(sn), _f32(tp), _f32(tn); sc, tc = _u8(sc), _u8(tc)
    m = np.ascontiguousarray(matches, MATCH_DTYPE).copy()
    lib().orc_apply_weights(C.c_int(method), C.c_float(max_d2), _p(sp), _p(tp), _p(sn), _p(tn), _p(sc), _p(tc), C.c_int64(len(m)), _p(m))
    return m


def prune(sn, tn, matches):
    """ICPOptimizer.h:157-174 (copy returned)."""
    sn, tn = _f32(sn), _f32(tn)
    m = np.ascontiguousarray(matches, MATCH_DTYPE).copy()
    lib().orc_prune(_p(sn), _p(tn), C.c_int64(len(m)), _p(m))
    return m


def match_pipeline(cfg: Config, pose, src, src_n, src_c, tgt, tgt_n, tgt_c, sel_idx=None, tree: KdTree | None = None,
                   return_transformed=False):
    """Stages 2-4 of one iteration at a given pose (teacher-forced parity)."""
    src, src_n, tgt, tgt_n = _f32(src), _f32(src_n), _f32(tgt), _f32(tgt_n)
    src_c, tgt_c = _u8(src_c), _u8(tgt_c)
    if tgt_c is None:
        tgt_c = np.zeros((len(tgt), 4), np.uint8)
    if tgt_n is None:
        tgt_n = np.zeros((len(tgt), 3), np.float32)
    sel = None if sel_idx is None else np.ascontiguousarray(sel_idx, dtype=np.int32)
    n_sel = len(src) if sel is None else len(sel)
    out = np.empty(n_sel, MATCH_DTYPE)
    tp = np.empty((n_sel, 3), np.float32); tn = np.empty((n_sel, 3), np.float32)
    cc = cfg.c()
    rc = lib().orc_match_pipeline(C.byref(cc), _p(_pose(pose)), _p(src), _p(src_n), _p(src_c), C.c_int64(len(src)),
                                  _p(sel), C.c_int64(n_sel), _p(tgt), _p(tgt_n), _p(tgt_c), C.c_int64(len(tgt)),
                                  tree._h if tree is not None else None, _p(out), _p(tp), _p(tn))
    if rc != 0:
        raise RuntimeError(f"orc_match_pipeline rc={rc}")
    return (out, tp, tn) if return_transformed else out


def solve_p2p(s, d, w):
    s, d = _f32(s), _f32(d); w = np.ascontiguousarray(w, np.float32); out = np.empty(16, np.float32)
    rc = lib().orc_solve_p2p(_p(s), _p(d), _p(w), C.c_int64(len(s)), _p(out))
    return rc, _unpose(out)


def solve_p2plane(s, d, n, w):
    s, d, n = _f32(s), _f32(d), _f32(n); w = np.ascontiguousarray(w, np.float32); out = np.empty(16, np.float32)
    rc = lib().orc_solve_p2plane(_p(s), _p(d), _p(n), _p(w), C.c_int64(len(s)), _p(out))
    return rc, _unpose(out)


def solve_symmetric(s, d, ns, nt, w):
    s, d, ns, nt = _f32(s), _f32(d), _f32(ns), _f32(nt); w = np.ascontiguousarray(w, np.float32); out = np.empty(16, np.float32)
    rc = lib().orc_solve_symmetric(_p(s), _p(d), _p(ns), _p(nt), _p(w), C.c_int64(len(s)), _p(out))
    return rc, _unpose(out)


def solve_lm(metric, sp, sn, tgt, tgt_n, matches, max_iterations=10):
    sp, sn, tgt, tgt_n = _f32(sp), _f32(sn), _f32(tgt), _f32(tgt_n)
    m = np.ascontiguousarray(matches, MATCH_DTYPE)
    x = np.zeros(6, np.float64); out = np.empty(16, np.float32); nlm = C.c_int(0)
    rc = lib().orc_solve_lm(C.c_int(metric), _p(sp), _p(sn), _p(tgt), _p(tgt_n), _p(m), C.c_int64(len(m)), C.c_int(max_iterations),
                            _p(x), _p(out), C.byref(nlm))
    return rc, x, _unpose(out), nlm.value


def estimate_pose(cfg: Config, src, src_n, src_c, tgt, tgt_n, tgt_c, init_pose=None):
    """Returns (rc, pose 4x4, pose history [iters,4,4], n_queries)."""
    src, src_n, tgt, tgt_n = _f32(src), _f32(src_n), _f32(tgt), _f32(tgt_n)
    src_c, tgt_c = _u8(src_c), _u8(tgt_c)
    if src_c is None:
        src_c = np.zeros((len(src), 4), np.uint8)
    if tgt_c is None:
        tgt_c = np.zeros((len(tgt), 4), np.uint8)
    pose = _pose(np.eye(4, dtype=np.float32) if init_pose is None else init_pose).copy()
    max_it = max(cfg.n_iterations, 64) + 64
    hist = np.zeros((max_it, 16), np.float32)
    n_it = C.c_int(0); nq = C.c_int64(0)
    cc = cfg.c()
    rc = lib().orc_estimate_pose(C.byref(cc), _p(src), _p(src_n), _p(src_c), C.c_int64(len(src)),
                                 _p(tgt), _p(tgt_n), _p(tgt_c), C.c_int64(len(tgt)), _p(pose), _p(hist), C.byref(n_it), C.byref(nq))
    h = np.stack([_unpose(hist[i]) for i in range(n_it.value)]) if n_it.value else np.zeros((0, 4, 4), np.float32)
    return rc, _unpose(pose), h, nq.value


def rmse(pose, src, ref):
    src, ref = _f32(src), _f32(ref)
    return float(lib().orc_rmse(_p(_pose(pose)), _p(src), _p(ref), C.c_int64(len(src))))


def pca_normals(pts, k=5, viewpoint=(0.0, 0.0, 0.0)):
    """PointCloud(pcl cloud) (PointCloud.h:41-76): (normals [N,3], curvature [N])."""
    pts = _f32(pts); nrm = np.empty_like(pts); cur = np.empty(len(pts), np.float32)
    vp = np.ascontiguousarray(viewpoint, np.float32)
    lib().orc_pca_normals(_p(pts), C.c_int64(len(pts)), C.c_int(k), _p(vp), _p(nrm), _p(cur))
    return nrm, cur


def benchmark_error(pose, src, ref):
    src, ref = _f32(src), _f32(ref)
    return float(lib().orc_benchmark_error(_p(_pose(pose)), _p(src), _p(ref), C.c_int64(len(src))))


def cloud_from_depth(depth, rgbx, fx, fy, cx, cy, extrinsics=None, keep_original_size=False, downsample=1, max_distance=0.1):
    """PointCloud(depthMap, colorFrame, ...) (PointCloud.h:78-165).  rgbx: flat uint8 RGBX frame (>= w*h + 3 bytes) or None."""
    depth = np.ascontiguousarray(depth, np.float32); h, w = depth.shape
    col = None if rgbx is None else np.ascontiguousarray(rgbx, np.uint8).reshape(-1)
    assert col is None or col.size >= h * w + 3
    K = np.ascontiguousarray(np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32).T.reshape(9))
    E = None if extrinsics is None else _pose(extrinsics)
    po = np.empty((h * w, 3), np.float32); no = np.empty((h * w, 3), np.float32); co = np.zeros((h * w, 4), np.uint8)
    n = lib().orc_cloud_from_depth(_p(depth), _p(col), _p(K), _p(E), C.c_uint32(w), C.c_uint32(h), C.c_int(int(keep_original_size)),
                                   C.c_uint32(downsample), C.c_float(max_distance), _p(po), _p(no), _p(co))
    return po[:n].copy(), no[:n].copy(), co[:n].copy()


def coarsest_stride(n):
    return int(lib().orc_coarsest_stride(C.c_int64(n)))


def voxel_indices(pts, nrm, stride):
    pts, nrm = _f32(pts), _f32(nrm)
    out = np.empty(len(pts), np.int32)
    lib().orc_voxel_indices.restype = C.c_int64
    c = lib().orc_voxel_indices(_p(pts), _p(nrm), C.c_int64(len(pts)), C.c_int(stride), _p(out))
    return out[:c].copy()


def coarse_indices(pts, nrm, stride):
    pts, nrm = _f32(pts), _f32(nrm)
    out = np.empty(len(pts), np.int32)
    c = lib().orc_coarse_indices(_p(pts), _p(nrm), C.c_int64(len(pts)), C.c_int(stride), _p(out))
    return out[:c].copy()


class MT19937(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]

    def seed(self, s):
        lib().orc_mt_seed(C.byref(self), C.c_uint32(s & 0xFFFFFFFF))
        return self

    def next(self):
        return int(lib().orc_mt_next(C.byref(self))) & 0xFFFFFFFF

    def canonical(self):
        return float(lib().orc_mt_canonical(C.byref(self)))


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))
