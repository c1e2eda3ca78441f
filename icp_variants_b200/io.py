"""The on-disk formats either side of the path, host side, dependency-free:

    OFF / COFF meshes       SimpleMesh::loadMesh / writeMesh          SimpleMesh.h:161-259
    TUM RGB-D lists         VirtualSensor::readFileList                VirtualSensor.h:196-216   (depth.txt / rgb.txt)
    TUM trajectory          VirtualSensor::readTrajectoryFile          VirtualSensor.h:218-250   (groundtruth.txt, inverted poses)
    TUM depth scaling       VirtualSensor::processFrameIndex           VirtualSensor.h:119-124   (u16 / 5000, 0 -> MINF)
    ETH pair list           ETHDataLoader::getItem + CSVReader         ETHDataLoader.h:50-61, CSVReader.h:27-44

    PCD point clouds        pcl::io::loadPCDFile<pcl::PointXYZ>        ETHDataLoader.h:68,87     (ascii and uncompressed binary)
    PLY point clouds        PointCloud::writeToFile / savePLYFile      PointCloud.h:219-236      (ascii and binary little-endian)
    PointCloud binary dump  PointCloud::readFromFile                   PointCloud.h:167-217

Decoding of PNG payloads (FreeImage in the reference) is left to the caller; these functions parse the formats and apply the
reference's conventions to already-decoded arrays."""
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


# --------------------------------------------------------------------------- PCD / PLY / PointCloud dump
_PCD_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4"}


def read_pcd(path: str):
    """pcl::io::loadPCDFile<pcl::PointXYZ> (ETHDataLoader.h:68,87): the x / y / z fields of an ascii or uncompressed-binary .pcd as
    [N,3] float32; NaN rows are kept (PCL keeps them too; the registration treats them as non-finite points).  Also returns a dict
    of the other fields (e.g. normal_x ..., intensity) as arrays."""
    with open(path, "rb") as f:
        raw = f.read()
    header, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        header[key.upper()] = val.split()
        if key.upper() == "DATA":
            break
    fields, sizes, types = header["FIELDS"], [int(v) for v in header["SIZE"]], header["TYPE"]
    counts = [int(v) for v in header.get("COUNT", ["1"] * len(fields))]
    n = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header.get("HEIGHT", ["1"])[0])
    mode = header["DATA"][0].lower()
    if mode == "ascii":
        tok = raw[pos:].split()
        per = sum(counts)
        a = np.array(tok[:n * per], dtype=np.float64).reshape(n, per)
        cols, k = {}, 0
        for name, c in zip(fields, counts):
            cols[name] = a[:, k] if c == 1 else a[:, k:k + c]
            k += c
    elif mode == "binary":
        dt = np.dtype([(name, _PCD_NP[(t.upper(), sz)], (c,) if c > 1 else ()) for name, t, sz, c in zip(fields, types, sizes, counts)])
        rec = np.frombuffer(raw, dtype=dt, count=n, offset=pos)
        cols = {name: rec[name] for name in fields}
    else:
        raise ValueError(f"unsupported PCD DATA mode {mode!r} (binary_compressed needs PCL's LZF)")
    xyz = np.stack([np.asarray(cols[k], np.float32) for k in ("x", "y", "z")], 1)
    return xyz, {k: np.asarray(v) for k, v in cols.items() if k not in ("x", "y", "z")}


def write_pcd(path: str, points, binary: bool = False):
    """An x y z .pcd (ascii or uncompressed binary) as PCL's PCDWriter lays it out (version 0.7 header)."""
    p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    hdr = (f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH {len(p)}\nHEIGHT 1\n"
           f"VIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(p)}\nDATA {'binary' if binary else 'ascii'}\n")
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        if binary:
            f.write(p.astype("<f4").tobytes())
        else:
            f.write("".join(f"{float(a)!r} {float(b)!r} {float(c)!r}\n" for a, b, c in p).encode("ascii"))


_PLY_NP = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
           "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


def read_ply(path: str):
    """The vertex element of a .ply (ascii, binary_little_endian or binary_big_endian) -- what PointCloud::writeToFile produces through
    pcl::io::savePLYFile (PointCloud.h:219-236: x y z intensity normal_x normal_y normal_z curvature).  Returns (points [N,3] float32,
    normals [N,3] float32 | None, dict of the remaining vertex properties)."""
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header")
    end = raw.index(b"\n", end) + 1
    lines = raw[:end].decode("ascii", "replace").splitlines()
    if not lines or lines[0].strip() != "ply":
        raise ValueError("not a PLY file")
    fmt, n, props, in_vertex, before = None, 0, [], False, 0
    for ln in lines[1:]:
        t = ln.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            in_vertex = t[1] == "vertex"
            if in_vertex:
                n = int(t[2])
            elif not props:
                before += 1          # elements before the vertex element are not supported (PCL and the reference write vertex first)
        elif t[0] == "property" and in_vertex:
            if t[1] == "list":
                raise ValueError("list properties in the vertex element are not supported")
            props.append((t[2], _PLY_NP[t[1]]))
    if before:
        raise ValueError("the vertex element must come first")
    if fmt == "ascii":
        tok = raw[end:].split()
        a = np.array(tok[:n * len(props)], dtype=np.float64).reshape(n, len(props))
        cols = {name: a[:, k] for k, (name, _) in enumerate(props)}
    else:
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(name, order + t if t[-1] != "1" else t) for name, t in props])
        rec = np.frombuffer(raw, dtype=dt, count=n, offset=end)
        cols = {name: rec[name] for name, _ in props}
    pts = np.stack([np.asarray(cols[k], np.float32) for k in ("x", "y", "z")], 1)
    nk = ("normal_x", "normal_y", "normal_z") if "normal_x" in cols else (("nx", "ny", "nz") if "nx" in cols else None)
    nrm = None if nk is None else np.stack([np.asarray(cols[k], np.float32) for k in nk], 1)
    used = {"x", "y", "z"} | set(nk or ())
    return pts, nrm, {k: np.asarray(v) for k, v in cols.items() if k not in used}


def write_ply(path: str, points, normals=None, binary: bool = False):
    """PointCloud::writeToFile (PointCloud.h:219-236): vertices with x y z [normal_x normal_y normal_z]; ascii or binary little-endian."""
    p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    cols = [p] + ([np.ascontiguousarray(normals, np.float32).reshape(-1, 3)] if normals is not None else [])
    a = np.concatenate(cols, 1)
    names = ["x", "y", "z"] + (["normal_x", "normal_y", "normal_z"] if normals is not None else [])
    hdr = "ply\nformat " + ("binary_little_endian" if binary else "ascii") + f" 1.0\nelement vertex {len(p)}\n" + \
          "".join(f"property float {k}\n" for k in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        if binary:
            f.write(a.astype("<f4").tobytes())
        else:
            f.write("".join(" ".join(repr(float(v)) for v in row) + "\n" for row in a).encode("ascii"))


def read_pointcloud_dump(path: str):
    """PointCloud::readFromFile (PointCloud.h:167-217): char nBytes (4 or 8), uint32 n, n points then n normals as float / double
    triples.  Returns (points, normals) as float32."""
    with open(path, "rb") as f:
        raw = f.read()
    nbytes = raw[0]
    n = int(np.frombuffer(raw, "<u4", 1, 1)[0])
    dt = "<f4" if nbytes == 4 else "<f8"
    a = np.frombuffer(raw, dt, 6 * n, 5).astype(np.float32)
    return a[:3 * n].reshape(n, 3).copy(), a[3 * n:].reshape(n, 3).copy()


def write_pointcloud_dump(path: str, points, normals, double: bool = False):
    p = np.ascontiguousarray(points).reshape(-1, 3); m = np.ascontiguousarray(normals).reshape(-1, 3)
    dt = "<f8" if double else "<f4"
    with open(path, "wb") as f:
        f.write(bytes([8 if double else 4])); f.write(np.uint32(len(p)).tobytes())
        f.write(p.astype(dt).tobytes()); f.write(m.astype(dt).tobytes())
