"""The on-disk formats either side of the path, host side, dependency-free:

    OFF / COFF meshes       SimpleMesh::loadMesh / writeMesh          SimpleMesh.h:161-259
    TUM RGB-D lists         VirtualSensor::readFileList                VirtualSensor.h:196-216   (depth.txt / rgb.txt)
    TUM trajectory          VirtualSensor::readTrajectoryFile          VirtualSensor.h:218-250   (groundtruth.txt, inverted poses)
    TUM depth scaling       VirtualSensor::processFrameIndex           VirtualSensor.h:119-124   (u16 / 5000, 0 -> MINF)
    ETH pair list           ETHDataLoader::getItem + CSVReader         ETHDataLoader.h:50-61, CSVReader.h:27-44

File decoding of PNG / PCD payloads (FreeImage, PCL in the reference) is left to the caller; these functions parse the
text formats and apply the reference's conventions to already-decoded arrays."""
from __future__ import annotations

import numpy as np

MINF = np.float32(-np.inf)


def read_off(path: str):
    """SimpleMesh::loadMesh (SimpleMesh.h:161-229).  Returns (vertices [N,3] float32, colors [N,4] uint8, faces [M,3] int32).
    'COFF' files carry integer RGBA per vertex; plain 'OFF' vertices get (0, 0, 0, 255) (:207-210).  Only triangles (:221)."""
    with open(path) as f:
        tok = f.read().split()
    kind = tok[0]
    if kind not in ("OFF", "COFF"):
        raise ValueError("Incorrect mesh file type.")                         # SimpleMesh.h:215-218
    nv, nf = int(tok[1]), int(tok[2])                                         # numV numP numE (numE unused)
    pos = 4
    per = 7 if kind == "COFF" else 3
    v = np.array(tok[pos:pos + per * nv], dtype=np.float64).reshape(nv, per)
    pos += per * nv
    vertices = v[:, :3].astype(np.float32)
    if kind == "COFF":
        colors = v[:, 3:7].astype(np.int64).astype(np.uint8)                  # (unsigned char)colorInt (:195)
    else:
        colors = np.tile(np.array([0, 0, 0, 255], np.uint8), (nv, 1))
    fa = np.array(tok[pos:pos + 4 * nf], dtype=np.int64).reshape(nf, 4)
    if nf and not (fa[:, 0] == 3).all():
        raise ValueError("We can only read triangular mesh.")                 # SimpleMesh.h:221
    return vertices, colors, fa[:, 1:4].astype(np.int32)


def write_off(path: str, vertices, colors=None, faces=None):
    """SimpleMesh::writeMesh (SimpleMesh.h:231-259): always 'COFF'; a vertex with a non-finite coordinate is written as
    '0.0 0.0 0.0 0 0 0 0' (:245-246)."""
    vertices = np.asarray(vertices, np.float32)
    n = len(vertices)
    colors = np.zeros((n, 4), np.uint8) if colors is None else np.asarray(colors, np.uint8)
    faces = np.zeros((0, 3), np.int32) if faces is None else np.asarray(faces, np.int32)
    with open(path, "w") as f:
        f.write("COFF\n")
        f.write(f"{n} {len(faces)} 0\n")
        for p, c in zip(vertices, colors):
            if np.isfinite(p).all():
                f.write(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r} {int(c[0])} {int(c[1])} {int(c[2])} {int(c[3])}\n")
            else:
                f.write("0.0 0.0 0.0 0 0 0 0\n")
        for t in faces:
            f.write(f"3 {int(t[0])} {int(t[1])} {int(t[2])}\n")


def read_tum_file_list(path: str):
    """VirtualSensor::readFileList (VirtualSensor.h:196-216): three header lines, then 'timestamp filename' records.
    Returns (timestamps float64 [n], filenames list)."""
    with open(path) as f:
        lines = f.read().split("\n")[3:]
    tok = " ".join(lines).split()
    ts = [float(tok[i]) for i in range(0, len(tok) - 1, 2)]
    names = [tok[i + 1] for i in range(0, len(tok) - 1, 2)]
    return np.array(ts, np.float64), names


def read_tum_trajectory(path: str):
    """VirtualSensor::readTrajectoryFile (VirtualSensor.h:218-250): 'timestamp tx ty tz qx qy qz qw' records after three
    header lines; each pose is stored INVERTED (world -> camera, :243); reading stops at a zero quaternion (:241).
    Returns (timestamps [n], poses [n,4,4] float32)."""
    with open(path) as f:
        tok = " ".join(f.read().split("\n")[3:]).split()
    ts, poses = [], []
    for i in range(0, len(tok) - 7, 8):
        t, tx, ty, tz, qx, qy, qz, qw = (float(x) for x in tok[i:i + 8])
        if qx * qx + qy * qy + qz * qz + qw * qw == 0.0:
            break
        # Eigen::Quaternionf::toRotationMatrix (no normalisation), fp32 like the reference
        x, y, z, w = (np.float32(v) for v in (qx, qy, qz, qw))
        tx2, ty2, tz2 = x + x, y + y, z + z
        r = np.array([[1 - (ty2 * y + tz2 * z), ty2 * x - tz2 * w, tz2 * x + ty2 * w],
                      [ty2 * x + tz2 * w, 1 - (tx2 * x + tz2 * z), tz2 * y - tx2 * w],
                      [tz2 * x - ty2 * w, tz2 * y + tx2 * w, 1 - (tx2 * x + ty2 * y)]], np.float32)
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = r
        m[:3, 3] = [tx, ty, tz]
        poses.append(np.linalg.inv(m.astype(np.float64)).astype(np.float32))
        ts.append(t)
    return np.array(ts, np.float64), (np.stack(poses) if poses else np.zeros((0, 4, 4), np.float32))


def nearest_trajectory_index(trajectory_timestamps, depth_timestamp: float) -> int:
    """VirtualSensor.h:126-137: the first trajectory record with the smallest |t - t_depth| (strict '>' scan)."""
    d = np.abs(np.asarray(trajectory_timestamps, np.float64) - float(depth_timestamp))
    return int(np.argmin(d)) if len(d) else 0


def tum_depth_to_float(depth_u16) -> np.ndarray:
    """VirtualSensor.h:119-124: metres = u16 * 1.0f / 5000.0f, 0 -> MINF."""
    d = np.asarray(depth_u16)
    out = (d.astype(np.float32) * np.float32(1.0)) / np.float32(5000.0)
    out[d == 0] = MINF
    return out.astype(np.float32)


def read_eth_pairs(path: str):
    """ETHDataLoader (ETHDataLoader.h:29-61) over CSVReader (CSVReader.h:27-44): the first row holds the column names;
    every other row = id, source file, target file, <unused>, then the 3x4 pose row-major in columns 4..15.
    Returns a list of dicts {id, source, target, pose [4,4] float32}."""
    out = []
    with open(path) as f:
        rows = [ln.rstrip("\r\n").split(",") for ln in f if ln.strip()]
    for vec in rows[1:]:
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :4] = np.array([float(x) for x in vec[4:16]], np.float32).reshape(3, 4)
        out.append({"id": vec[0], "source": vec[1], "target": vec[2], "pose": pose})
    return out
