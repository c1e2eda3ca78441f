"""Oracle stages 1,3,4 + projective matching against independent numpy restatements and the
reference's documented quirks."""
import numpy as np

from oracle import oracle as O


def test_mt19937_matches_std_and_numpy():
    # std::mt19937 default-seeded: 10000th output is 4123659995 (C++ standard, [rand.predef])
    r = O.MT19937().seed(5489)
    v = 0
    for _ in range(10000):
        v = r.next()
    assert v == 4123659995
    # same seeding (init_genrand) and tempering as numpy's legacy RandomState
    r = O.MT19937().seed(1234)
    rs = np.random.RandomState(1234)
    ref = rs.randint(0, 2 ** 32, size=50, dtype=np.uint64)
    assert [r.next() for _ in range(50)] == ref.tolist()


def test_generate_canonical_two_draws():
    r = O.MT19937().seed(42)
    r2 = O.MT19937().seed(42)
    for _ in range(100):
        lo, hi = r2.next(), r2.next()
        assert r.canonical() == (lo + hi * 2.0 ** 32) / 2.0 ** 64


def test_coarsest_stride_table():
    # SURVEY Appendix A / ICPOptimizer.h:503-516
    assert O.coarsest_stride(1054) == 8
    assert O.coarsest_stride(38400) == 256
    assert O.coarsest_stride(307200) == 2048
    assert O.coarsest_stride(370488) == 2048
    assert O.coarsest_stride(3000000) == 16384


def test_coarse_indices_filters_non_finite():
    p = np.arange(30, dtype=np.float32).reshape(10, 3)
    n = np.ones((10, 3), np.float32)
    p[4, 1] = -np.inf
    n[6, 0] = np.nan
    assert O.coarse_indices(p, n, 2).tolist() == [0, 2, 8]
    assert O.coarse_indices(p, n, 1).tolist() == [0, 1, 2, 3, 5, 7, 8, 9]


def test_transform_matches_numpy():
    rng = np.random.default_rng(0)
    from icp_variants_b200.synth import make_pose
    pose = make_pose([0.3, -0.2, 0.1], [10, -20, 30])
    p = rng.normal(size=(1000, 3)).astype(np.float32)
    ref = p.astype(np.float64) @ pose[:3, :3].astype(np.float64).T + pose[:3, 3]
    assert np.allclose(O.transform_points(pose, p), ref, atol=1e-6)
    refn = p.astype(np.float64) @ np.linalg.inv(pose[:3, :3].astype(np.float64))  # (R^-1)^T n == n^T R^-1
    assert np.allclose(O.transform_normals(pose, p), refn, atol=2e-6)


def _pipeline(cfg, sp, sn, sc, tp, tn, tc):
    return O.match_pipeline(cfg, np.eye(4, dtype=np.float32), sp, sn, sc, tp, tn, tc)


def test_weighting_and_rejection_rules():
    tp = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    tn = np.array([[0, 0, 1], [0, 0, 1]], np.float32)
    tc = np.array([[10, 20, 30, 0], [250, 0, 0, 0]], np.uint8)
    sp = np.array([[0.1, 0, 0], [1, 0.2, 0], [0, 0, 0.05], [1, 0, 0]], np.float32)
    c60 = np.float32(0.5)
    sn = np.array([[0, 0, 1], [0, np.sqrt(1 - 0.25), 0.5], [0, 0, -1], [0, np.sqrt(np.float32(1) - np.nextafter(c60, np.float32(1)) ** 2), np.nextafter(c60, np.float32(1))]], np.float32)
    sc = np.array([[12, 20, 30, 9], [5, 0, 0, 9], [10, 20, 30, 9], [250, 0, 0, 9]], np.uint8)
    # distance weighting (weighting.h:16-20)
    m = _pipeline(O.Config(weighting=1, rejection=0, max_distance_sq=0.25), sp, sn, sc, tp, tn, tc)
    assert m["idx"].tolist() == [0, 1, 0, 1]
    exp = [np.float32(1.0 - np.float64(np.float32(np.float32(d2) / np.float32(0.25)))) for d2 in
           (np.float32(0.1) * np.float32(0.1), np.float32(0.2) * np.float32(0.2), np.float32(0.05) * np.float32(0.05), 0.0)]
    assert m["weight"].tolist() == [float(e) for e in exp]
    # normals weighting is the raw dot product, may be negative (weighting.h:22-25)
    m = _pipeline(O.Config(weighting=2, rejection=0, max_distance_sq=0.25), sp, sn, sc, tp, tn, tc)
    assert m["weight"][0] == 1.0 and m["weight"][2] == -1.0
    # colour weighting: uchar difference wraps (5-250 -> 11), weighting.h:27-30
    m = _pipeline(O.Config(weighting=3, rejection=0, max_distance_sq=0.25), sp, sn, sc, tp, tn, tc)
    w_dist1 = np.float32(1.0 - np.float64(np.float32(np.float32(np.float32(0.2) * np.float32(0.2)) / np.float32(0.25))))
    w_col1 = np.float32(1.0 - np.float64(np.float32(11 * 11) / np.float32(195075)))
    assert m["weight"][1] == np.float32(w_dist1 * w_col1)
    w_col0 = np.float32(1.0 - np.float64(np.float32(4) / np.float32(195075)))
    assert m["weight"][0] == np.float32(exp[0] * w_col0)
    # rejection (ICPOptimizer.h:157-174): exactly-60-degree normals (cos == 0.5f) are rejected,
    # the next float above 0.5 is kept, opposite normals rejected; weights are left untouched.
    m = _pipeline(O.Config(weighting=0, rejection=1, max_distance_sq=0.25), sp, sn, sc, tp, tn, tc)
    assert m["idx"].tolist()[0] == 0 and m["idx"][2] == -1
    assert m["weight"].tolist() == [1.0, 1.0, 1.0, 1.0]
    cos1 = (sn[1] * tn[1]).sum() / np.linalg.norm(sn[1])
    assert (m["idx"][1] == -1) == (np.float32(cos1) <= np.float32(0.5))


def test_rejection_boundary_is_cos_le_half():
    # acosf(0.5f) = 1.04719758 > 60*pi/180 = 1.0471975512 -> rejected ; nextafter(0.5f,1) kept
    tp = np.zeros((1, 3), np.float32)
    tn = np.array([[0, 0, 1]], np.float32)
    for c, rejected in ((np.float32(0.5), True), (np.nextafter(np.float32(0.5), np.float32(1)), False),
                        (np.nextafter(np.float32(0.5), np.float32(0)), True)):
        # a source normal of length 1/c along z' so that dot/(|a||b|) is exactly representable: use (0,0,1) vs target (s,0,c)
        s = np.float32(np.sqrt(np.float64(1) - np.float64(c) ** 2))
        tn2 = np.array([[s, 0, c]], np.float32)
        m = _pipeline(O.Config(rejection=1, max_distance_sq=1.0), tp, tn, None, tp, tn2, None)
        nb = np.sqrt(np.float32(np.float32(s * s) + np.float32(c * c)))
        cosv = np.float32(c / np.float32(np.float32(1.0) * nb))
        assert (m["idx"][0] == -1) == bool(cosv <= np.float32(0.5)), (c, cosv)
        if cosv == c:
            assert (m["idx"][0] == -1) == rejected


def _numpy_projective(tgt, w, h, fx, fy, cx, cy, q, max_d2, win=12):
    out = []
    for p in q:
        if p[0] == -np.inf:
            out.append((0, 0.0)); continue
        with np.errstate(all="ignore"):
            uf = np.round(np.float32(np.float32(p[0] * np.float32(fx)) / p[2]) + np.float32(cx))
            vf = np.round(np.float32(np.float32(p[1] * np.float32(fy)) / p[2]) + np.float32(cy))
        def conv(t):
            if not np.isfinite(t) or abs(t) >= 2.0 ** 63:
                return 0
            return int(t) % (1 << 32)
        u0, v0 = conv(uf), conv(vf)
        best, bi = np.float32(np.finfo(np.float32).max), -1
        if u0 >= win and v0 >= win:
            for v in range(v0 - win, min(v0 + win, h - 1) + 1):
                for u in range(u0 - win, min(u0 + win, w - 1) + 1):
                    t = tgt[v * w + u]
                    if t[0] == -np.inf:
                        continue
                    d = p - t
                    d2 = np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
                    if best > d2:
                        best, bi = d2, v * w + u
        out.append((bi, 1.0) if best <= max_d2 else (-1, 0.0))
    return out


def test_projective_quirks():
    rng = np.random.default_rng(5)
    w, h, fx, fy, cx, cy = 64, 48, 52.5, 52.5, 31.5, 23.5
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    z = (2.0 + 0.3 * np.sin(u / 7.0) + 0.2 * np.cos(v / 5.0)).astype(np.float32)
    tgt = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], -1).reshape(-1, 3).astype(np.float32)
    tgt[rng.random(len(tgt)) < 0.1] = -np.inf
    q = tgt + rng.normal(0, 0.01, size=tgt.shape).astype(np.float32)
    q = np.concatenate([q, np.array([[0, 0, 0], [1, 1, -1], [-5, 0, 1], [0, -5, 1], [np.nan, 0, 1], [1e9, 0, 1e-9]], np.float32)])
    got = O.projective(tgt, w, h, fx, fy, cx, cy, q, 0.01)
    ref = _numpy_projective(tgt, w, h, fx, fy, cx, cy, q, np.float32(0.01))
    assert got["idx"].tolist() == [r[0] for r in ref]
    assert got["weight"].tolist() == [r[1] for r in ref]
    # the low-border quirk: pixels projecting to u<12 or v<12 never match (unsigned wrap, NearestNeighbor.h:385-386)
    uu = (np.arange(len(tgt)) % w)
    vv = (np.arange(len(tgt)) // w)
    inner = got["idx"][:len(tgt)]
    assert (inner[(uu < 11) | (vv < 11)] <= 0).all()
    # invalid source pixel keeps the value-initialised Match{0, 0.f} (NearestNeighbor.h:353,372-373)
    bad = np.where(q[:len(tgt), 0] == -np.inf)[0]
    assert (got["idx"][bad] == 0).all() and (got["weight"][bad] == 0).all()


def test_voxel_levels_properties(small_eth_pair):
    """Voxel pyramid levels (extension): every level is a subset of the valid points, holds the lowest index of each
    occupied cell, is nested (a coarser level's representatives survive in the finer ones) and grows with depth."""
    import numpy as np
    from oracle import oracle as orc
    src, _, _ = small_eth_pair
    pts, nrm = src.points.copy(), src.normals.copy()
    nrm[3::50] = np.nan
    valid = np.isfinite(pts).all(1) & np.isfinite(nrm).all(1)
    prev = None
    for stride in (64, 32, 16, 8, 4, 2):
        lv = orc.voxel_indices(pts, nrm, stride)
        assert np.all(np.diff(lv) > 0) and valid[lv].all()
        if prev is not None:
            assert len(lv) >= len(prev) and np.isin(prev, lv).all()
        prev = lv
    assert np.array_equal(orc.voxel_indices(pts, nrm, 1), np.nonzero(valid)[0])
    assert 0 in orc.voxel_indices(pts, nrm, 64) or not valid[0]     # the lowest index always represents its cell
