"""Input preparation for the registration path: the bundled bunny meshes and synthetic clouds with
the shapes BASELINE.json names (ETH-Apartment-like tilting-lidar scans, TUM freiburg1-like
640x480 RGB-D frames).  numpy only; this is harness-side data preparation (the reference does the
same work on the host in PointCloud.h / *DataLoader.h), not part of the timed path.

Conventions: points / normals float32 [N,3], colours uint8 [N,4], poses 4x4 float32 row-major
numpy (converted to Eigen's column-major float[16] at the C ABI).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

MINF = np.float32(-np.inf)
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


@dataclass
class Cloud:
    """PointCloud (PointCloud.h): AoS points / normals / colours."""
    points: np.ndarray
    normals: np.ndarray
    colors: np.ndarray

    def __len__(self):
        return len(self.points)


# --------------------------------------------------------------------------- bunny (config C1)

def mesh_to_cloud(vertices: np.ndarray, faces: np.ndarray) -> Cloud:
    """PointCloud(const SimpleMesh&) (PointCloud.h:12-39): normals = normalised sum of the
    (un-normalised, i.e. area-weighted) face normals; colours zeroed (:26)."""
    v = np.ascontiguousarray(vertices, np.float32)
    n = np.zeros_like(v)
    for i0, i1, i2 in np.asarray(faces):
        e1 = v[i1] - v[i0]
        e2 = v[i2] - v[i0]
        fn = np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]], np.float32)
        n[i0] += fn
        n[i1] += fn
        n[i2] += fn
    z = ((n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]) + n[:, 2] * n[:, 2]).astype(np.float32)
    nz = z > 0
    n[nz] = (n[nz] / np.sqrt(z[nz])[:, None]).astype(np.float32)
    return Cloud(v, n, np.zeros((len(v), 4), np.uint8))


def load_bunny(path: str | None = None):
    """BunnyDataLoader (BunnyDataLoader.h:10-11,34-37): source = bunny_part2_trans, target = bunny_part1.
    Returns (source Cloud, target Cloud, gt_source_idx, gt_target_idx)."""
    z = np.load(path or os.path.join(_GOLDEN, "bunny.npz"))
    src = mesh_to_cloud(z["source_vertices"], z["source_faces"])
    tgt = mesh_to_cloud(z["target_vertices"], z["target_faces"])
    return src, tgt, z["gt_source_idx"], z["gt_target_idx"]


# --------------------------------------------------------------------------- helpers

def rot_xyz(ax, ay, az):
    """Rx(ax) * Ry(ay) * Rz(az) (ICPOptimizer.h:771-773, main.cpp:421-424)."""
    ca, sa, cb, sb, cg, sg = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
    return rx @ ry @ rz


def make_pose(t, angles_deg):
    p = np.eye(4)
    p[:3, :3] = rot_xyz(*np.deg2rad(angles_deg))
    p[:3, 3] = t
    return p.astype(np.float32)


def apply_pose(pose, cloud: Cloud) -> Cloud:
    """PointCloud::change_pose (PointCloud.h:263-268)."""
    r = pose[:3, :3].astype(np.float32)
    t = pose[:3, 3].astype(np.float32)
    return Cloud((cloud.points @ r.T + t).astype(np.float32), (cloud.normals @ r.T).astype(np.float32), cloud.colors.copy())


# --------------------------------------------------------------------------- the synthetic room

@dataclass
class Room:
    size: np.ndarray          # (3,) box room [0,size]
    boxes_lo: np.ndarray      # (B,3)
    boxes_hi: np.ndarray      # (B,3)


def make_room(seed=1234, size=(17.0, 10.0, 3.0), n_boxes=10) -> Room:
    rng = np.random.default_rng(seed)
    size = np.asarray(size, np.float64)
    lo, hi = [], []
    for _ in range(n_boxes):
        ext = np.array([rng.uniform(0.5, 2.5), rng.uniform(0.5, 2.0), rng.uniform(0.4, 2.0)])
        base = np.array([rng.uniform(0.3, size[0] - ext[0] - 0.3), rng.uniform(0.3, size[1] - ext[1] - 0.3), 0.0])
        lo.append(base)
        hi.append(base + ext)
    return Room(size, np.array(lo), np.array(hi))


def raycast(room: Room, origin, dirs):
    """Nearest hit distance along unit rays from a point inside the room. dirs [N,3] float64."""
    o = np.asarray(origin, np.float64)
    d = np.asarray(dirs, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # room walls: exit distance of the enclosing box
        t1 = (0.0 - o) * inv
        t2 = (room.size - o) * inv
        t_exit = np.min(np.maximum(t1, t2), axis=1)
        best = t_exit
        for lo, hi in zip(room.boxes_lo, room.boxes_hi):
            a = (lo - o) * inv
            b = (hi - o) * inv
            tn = np.max(np.minimum(a, b), axis=1)
            tf = np.min(np.maximum(a, b), axis=1)
            hit = (tn < tf) & (tn > 1e-6)
            best = np.where(hit & (tn < best), tn, best)
    return best


def pca_normals(points: np.ndarray, viewpoint, k=5) -> np.ndarray:
    """PCL NormalEstimation semantics used by PointCloud(pcl cloud) (PointCloud.h:41-56): k nearest
    neighbours (the point itself included), normal = eigenvector of the smallest eigenvalue of
    their covariance, flipped towards the viewpoint."""
    from scipy.spatial import cKDTree
    p64 = points.astype(np.float64)
    _, nb = cKDTree(p64).query(p64, k=k)
    nbr = p64[nb]                                   # [N,k,3]
    c = nbr - nbr.mean(axis=1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / k
    _, vec = np.linalg.eigh(cov)
    n = vec[:, :, 0]
    flip = np.einsum("ni,ni->n", n, np.asarray(viewpoint, np.float64) - p64) < 0
    n[flip] *= -1.0
    return n.astype(np.float32)


def lidar_scan(room: Room, sensor_pos, yaw_deg=0.0, n_sweeps=344, n_beams=1077, noise=0.01, max_range=30.0,
               seed=0, normals_k=5, colors=None, normals_fn=None) -> Cloud:
    """ETH 'Challenging data sets'-like tilting 2-D lidar: n_sweeps tilt steps (-45..+45 deg) of a
    270-degree, n_beams-beam planar scan (README.md:45-46: ~370k points per scan), range noise
    N(0, noise), expressed in the world frame.  Normals: k=5 PCA towards the sensor; colours
    (255,255,255,1) as PointCloud.h:74 unless a procedural texture is requested."""
    rng = np.random.default_rng(seed)
    az = np.deg2rad(np.linspace(-135.0, 135.0, n_beams)) + np.deg2rad(yaw_deg)
    tilt = np.deg2rad(np.linspace(-45.0, 45.0, n_sweeps))
    azg, tg = np.meshgrid(az, tilt)                 # [sweeps, beams]
    # planar scan in the sensor x-y plane, plane tilted about the sensor's y axis
    dx = np.cos(azg) * np.cos(tg)
    dy = np.sin(azg)
    dz = np.cos(azg) * np.sin(tg)
    dirs = np.stack([dx, dy, dz], -1).reshape(-1, 3)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    t = raycast(room, sensor_pos, dirs)
    t = t + rng.normal(0.0, noise, size=t.shape)
    keep = (t > 0.05) & (t < max_range)
    pts = (np.asarray(sensor_pos, np.float64) + dirs[keep] * t[keep, None]).astype(np.float32)
    # normals_fn(points, viewpoint) -> normals: e.g. the device's k = 5 PCA normals (seconds -> milliseconds for big scans)
    nrm = normals_fn(pts, np.asarray(sensor_pos, np.float32)) if normals_fn is not None else pca_normals(pts, sensor_pos, k=normals_k)
    if colors == "texture":
        col = procedural_colors(pts)
    else:
        col = np.tile(np.array([255, 255, 255, 1], np.uint8), (len(pts), 1))
    return Cloud(pts, nrm, col)


def procedural_colors(points: np.ndarray, seed=99) -> np.ndarray:
    """3-D checker + low-frequency variation (config C4 texture)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, size=(3, 3))
    p = points.astype(np.float64)
    chk = (np.floor(p[:, 0] / 0.5) + np.floor(p[:, 1] / 0.5) + np.floor(p[:, 2] / 0.5)) % 2
    col = np.zeros((len(p), 4), np.uint8)
    for c in range(3):
        low = 0.5 + 0.5 * np.sin(0.7 * p[:, 0] + ph[c, 0]) * np.sin(0.9 * p[:, 1] + ph[c, 1]) * np.sin(1.3 * p[:, 2] + ph[c, 2])
        col[:, c] = np.clip(60 + 120 * chk + 70 * low, 0, 255).astype(np.uint8)
    col[:, 3] = 255
    return col


def eth_pair(seed=1234, n_sweeps=344, n_beams=1077, noise=0.01, pose_scaling=0.1, colors=None, pair_index=0, normals_fn=None):
    """One ETH-Apartment-shaped scan pair as alignETH prepares it (main.cpp:411-429): both scans in a
    common frame, the source then moved by the ground-truth perturbation scaled by 0.1.
    Returns (source Cloud, target Cloud, applied perturbation 4x4)."""
    room = make_room(seed)
    rng = np.random.default_rng(seed + 7919 * (pair_index + 1))
    base = np.array([6.0, 4.5, 1.2]) + np.array([0.25 * pair_index, 0.05 * pair_index, 0.0])
    base = np.clip(base, [1.0, 1.0, 0.8], room.size - [1.0, 1.0, 0.8])
    p0 = base
    p1 = base + np.array([0.30, 0.20, 0.05])
    tgt = lidar_scan(room, p0, yaw_deg=3.0 * pair_index, n_sweeps=n_sweeps, n_beams=n_beams, noise=noise,
                     seed=int(rng.integers(1 << 30)), colors=colors, normals_fn=normals_fn)
    src = lidar_scan(room, p1, yaw_deg=3.0 * pair_index + 5.0, n_sweeps=n_sweeps, n_beams=n_beams, noise=noise,
                     seed=int(rng.integers(1 << 30)), colors=colors, normals_fn=normals_fn)
    full_t = np.array([0.30, 0.20, 0.05])
    full_a = np.array([2.0, 1.0, 5.0])
    pert = make_pose(pose_scaling * full_t, pose_scaling * full_a)
    return apply_pose(pert, src), tgt, pert


# --------------------------------------------------------------------------- TUM-shaped RGB-D frames (config C3)

TUM_FX, TUM_FY, TUM_CX, TUM_CY, TUM_W, TUM_H = 525.0, 525.0, 319.5, 239.5, 640, 480   # VirtualSensor.h:38-46


def render_depth(room: Room, cam_pos, cam_yaw_deg=0.0, cam_pitch_deg=0.0, width=TUM_W, height=TUM_H, fx=TUM_FX, fy=TUM_FY,
                 cx=TUM_CX, cy=TUM_CY, seed=0, dropout=0.15):
    """Pin-hole depth image of the room: z-depth quantised to 1/5000 m (VirtualSensor.h:119-124),
    invalid pixels = MINF: blob dropouts covering ~`dropout` of the image plus depth-jump pixels."""
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    d_cam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1).reshape(-1, 3)
    # camera looks along world +x, image x -> world -y, image y -> world -z, then yaw/pitch
    base = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    r = rot_xyz(0.0, np.deg2rad(cam_pitch_deg), np.deg2rad(cam_yaw_deg)) @ base
    d_w = d_cam @ r.T
    norm = np.linalg.norm(d_w, axis=1)
    t = raycast(room, cam_pos, d_w / norm[:, None])
    z = (t / norm).reshape(height, width)
    z = np.round(z * 5000.0) / 5000.0
    depth = z.astype(np.float32)
    # blob dropouts
    mask = np.zeros((height, width), bool)
    n_blobs = int(dropout * width * height / (np.pi * 20.0 ** 2)) if dropout > 0 else 0
    yy, xx = np.mgrid[0:height, 0:width]
    for _ in range(n_blobs):
        bx, by, br = rng.uniform(0, width), rng.uniform(0, height), rng.uniform(8, 28)
        mask |= (xx - bx) ** 2 + (yy - by) ** 2 < br * br
    jump = np.zeros_like(mask)
    jump[:, 1:] |= np.abs(np.diff(z, axis=1)) > 0.15
    jump[1:, :] |= np.abs(np.diff(z, axis=0)) > 0.15
    depth[mask | jump] = MINF
    return depth, r.astype(np.float32)


def depth_to_cloud(depth: np.ndarray, colors: np.ndarray | None = None, fx=TUM_FX, fy=TUM_FY, cx=TUM_CX, cy=TUM_CY,
                   keep_original_size=False, downsample=1, max_distance=0.1) -> Cloud:
    """PointCloud(depthMap, colorFrame, intrinsics, extrinsics=I, w, h, keepOriginalSize, downsampleFactor,
    maxDistance) (PointCloud.h:78-165) with identity extrinsics, fp32 like the reference."""
    h, w = depth.shape
    f32 = np.float32
    u, v = np.meshgrid(np.arange(w, dtype=f32), np.arange(h, dtype=f32))
    d = depth.astype(f32)
    valid = d != MINF
    with np.errstate(invalid="ignore"):
        x = ((u - f32(cx)) / f32(fx) * d).astype(f32)
        y = ((v - f32(cy)) / f32(fy) * d).astype(f32)
    pts = np.stack([x, y, d], -1).astype(f32)
    pts[~valid] = MINF
    nrm = np.full((h, w, 3), MINF, f32)
    half = f32(max_distance / 2.0)
    with np.errstate(invalid="ignore"):
        du = f32(0.5) * (d[1:-1, 2:] - d[1:-1, :-2])
        dv = f32(0.5) * (d[2:, 1:-1] - d[:-2, 1:-1])
    ok = np.isfinite(du) & np.isfinite(dv) & (np.abs(du) <= half) & (np.abs(dv) <= half)
    nn = np.stack([-du, -dv, np.ones_like(du)], -1).astype(f32)
    sq = ((nn[..., 0] * nn[..., 0] + nn[..., 1] * nn[..., 1]) + nn[..., 2] * nn[..., 2]).astype(f32)
    nn = (nn / np.sqrt(sq)[..., None]).astype(f32)
    inner = nrm[1:-1, 1:-1]
    inner[ok] = nn[ok]
    pts = pts.reshape(-1, 3)
    nrm = nrm.reshape(-1, 3)
    if colors is None:
        colors = np.zeros((h * w, 4), np.uint8)
    colors = colors.reshape(-1, 4)
    idx = np.arange(0, h * w, downsample)
    if not keep_original_size:
        fin = np.isfinite(pts[idx]).all(1) & np.isfinite(nrm[idx]).all(1)
        idx = idx[fin]
    return Cloud(np.ascontiguousarray(pts[idx]), np.ascontiguousarray(nrm[idx]), np.ascontiguousarray(colors[idx]))


def tum_pair(seed=1234, frame_gap=10, width=TUM_W, height=TUM_H, dropout=0.15):
    """Two TUM-shaped frames `frame_gap` frames apart (1.5 mm / 0.1 deg per frame), both as full-size
    camera-frame clouds (keepOriginalSize=true, as reconstructRoom builds them for projective matching
    / multires, main.cpp:202-206,295-298).  Returns (source Cloud, target Cloud, K 3x3, gt pose
    source->target 4x4)."""
    room = make_room(seed)
    sx = width / TUM_W
    fx, fy, cx, cy = TUM_FX * sx, TUM_FY * sx, (TUM_CX + 0.5) * sx - 0.5, (TUM_CY + 0.5) * sx - 0.5
    pos0 = np.array([3.0, 5.0, 1.4])
    pos1 = pos0 + frame_gap * np.array([0.0015, 0.0006, 0.0002])
    d0, r0 = render_depth(room, pos0, 10.0, 0.0, width, height, fx, fy, cx, cy, seed=seed, dropout=dropout)
    d1, r1 = render_depth(room, pos1, 10.0 + 0.1 * frame_gap, 0.0, width, height, fx, fy, cx, cy, seed=seed + 1, dropout=dropout)
    tgt = depth_to_cloud(d0, None, fx, fy, cx, cy, keep_original_size=True)
    src = depth_to_cloud(d1, None, fx, fy, cx, cy, keep_original_size=True)
    k = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    # camera-to-world of each frame: x_w = r * x_c + pos ;  source->target = T0^-1 * T1
    t0 = np.eye(4); t0[:3, :3] = r0; t0[:3, 3] = pos0
    t1 = np.eye(4); t1[:3, :3] = r1; t1[:3, 3] = pos1
    gt = np.linalg.inv(t0) @ t1
    return src, tgt, k, gt.astype(np.float32)


def tum_sequence(n_frames=11, seed=1234, frame_step=10, width=TUM_W, height=TUM_H, dropout=0.15):
    """A TUM-freiburg1-shaped RGB-D sequence as reconstructRoom replays it (main.cpp:183-341): frame 0 plus every
    `frame_step`-th frame (1.5 mm / 0.1 deg of motion per frame).  Returns (depth frames [n,h,w] float32 with MINF holes,
    K 3x3, gt poses [n,4,4] mapping frame i's camera space to frame 0's = targetTrajectory * trajectory_i^-1)."""
    room = make_room(seed)
    sx = width / TUM_W
    fx, fy, cx, cy = TUM_FX * sx, TUM_FY * sx, (TUM_CX + 0.5) * sx - 0.5, (TUM_CY + 0.5) * sx - 0.5
    pos0 = np.array([3.0, 5.0, 1.4])
    frames, cams = [], []
    for i in range(n_frames):
        k = i * frame_step
        pos = pos0 + k * np.array([0.0015, 0.0006, 0.0002])
        d, r = render_depth(room, pos, 10.0 + 0.1 * k, 0.0, width, height, fx, fy, cx, cy, seed=seed + i, dropout=dropout)
        t = np.eye(4); t[:3, :3] = r; t[:3, 3] = pos
        frames.append(d); cams.append(t)
    gt = np.stack([np.linalg.inv(cams[0]) @ c for c in cams]).astype(np.float32)
    k3 = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    return np.stack(frames), k3, gt
