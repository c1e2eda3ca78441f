#!/usr/bin/env python3
"""Regenerates tests/golden/bunny.npz from the reference's bundled OFF/COFF meshes.

Input data (NOT code): /root/reference/Data/bunny_part1.off (target, COFF) and
bunny_part2_trans.off (source).  Parsing follows the reference reader
(icp-variants/SimpleMesh.h:161-229): header token, `numV numP numE`, then per vertex
`x y z [r g b a]` read with `istream >> float` (decimal -> nearest float32), then per
face `3 i0 i1 i2`.  Only runs in the build container (the GPU box has no /root/reference);
the .npz it writes is the committed fixture.
"""
import sys
import numpy as np

REF = "/root/reference/Data"


def read_off(path):
    tok = open(path).read().split()
    kind = tok[0]
    assert kind in ("OFF", "COFF"), kind
    nv, nf = int(tok[1]), int(tok[2])
    pos = 4
    per = 7 if kind == "COFF" else 3
    v = np.array(tok[pos:pos + per * nv], dtype=np.float64).reshape(nv, per)
    pos += per * nv
    f = np.array(tok[pos:pos + 4 * nf], dtype=np.int64).reshape(nf, 4)
    assert (f[:, 0] == 3).all()
    verts = np.array([[np.float32(s) for s in tok[4 + per * i: 4 + per * i + 3]] for i in range(nv)], dtype=np.float32)
    cols = v[:, 3:7].astype(np.uint8) if kind == "COFF" else np.tile(np.array([0, 0, 0, 255], np.uint8), (nv, 1))
    return verts, f[:, 1:4].astype(np.int32), cols


def main():
    tv, tf, tc = read_off(f"{REF}/bunny_part1.off")
    sv, sf, sc = read_off(f"{REF}/bunny_part2_trans.off")
    assert tv.shape == (1359, 3) and tf.shape == (2575, 3)
    assert sv.shape == (1054, 3) and sf.shape == (2002, 3)
    out = sys.argv[1] if len(sys.argv) > 1 else "tests/golden/bunny.npz"
    np.savez_compressed(out, target_vertices=tv, target_faces=tf, target_colors=tc,
                        source_vertices=sv, source_faces=sf, source_colors=sc,
                        # ground-truth correspondences hard-coded in the reference drivers
                        # (icp-variants/main.cpp:105-120, experiment.cpp:64-79)
                        gt_source_idx=np.array([215, 424, 640, 1023], np.int32),
                        gt_target_idx=np.array([294, 258, 1238, 1310], np.int32))
    print("wrote", out)


if __name__ == "__main__":
    main()
