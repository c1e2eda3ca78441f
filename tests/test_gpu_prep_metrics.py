"""GPU parity of the steps either side of the loop, through the C ABI:
depth map -> cloud (PointCloud.h:78-165) bit-exact against the oracle (itself pinned against the reference's
constructor in tests/test_oracle_vs_reference.py), and the per-iteration convergence metrics
(ConvergenceMeasure.h:50-66, :104-151) evaluated on the device."""
import numpy as np
import pytest

from icp_variants_b200 import capi, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _frame(w, h, seed):
    fx = fy = 525.0 * w / 640; cx, cy = (319.5 + 0.5) * w / 640 - 0.5, (239.5 + 0.5) * w / 640 - 0.5
    depth, _ = synth.render_depth(synth.make_room(seed), np.array([3.0, 5.0, 1.4]), 10.0, 0.0, w, h, fx, fy, cx, cy, seed=seed)
    rgbx = np.random.default_rng(seed).integers(0, 256, 4 * h * w, dtype=np.uint8)
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    return depth, rgbx, K


@pytest.mark.parametrize("w,h", [(96, 72), (640, 480), (33, 17)])
def test_cloud_from_depth_bit_exact(ctx, w, h):
    depth, rgbx, K = _frame(w, h, 3)
    E = synth.make_pose([0.1, -0.2, 0.05], [5, -3, 10])
    for keep, ds, md, ext in ((True, 1, 0.1, None), (False, 1, 0.1, None), (False, 8, 0.1, None), (False, 3, 0.05, None), (False, 1, 0.1, E), (True, 5, 0.1, E)):
        po, no, co = orc.cloud_from_depth(depth, rgbx, K[0, 0], K[1, 1], K[0, 2], K[1, 2], ext, keep, ds, md)
        pg, ng, cg = ctx.cloud_from_depth(depth, rgbx, K, ext, keep, ds, md)
        assert pg.shape == po.shape, (keep, ds, md)
        assert np.array_equal(pg, po, equal_nan=True) and np.array_equal(ng, no, equal_nan=True) and np.array_equal(cg, co)
    pg, ng, cg = ctx.cloud_from_depth(depth, None, K, None, True, 1, 0.1)
    assert (cg == 0).all() and len(pg) == w * h


def test_cloud_from_depth_empty_and_errors(ctx):
    depth = np.full((8, 8), -np.inf, np.float32)
    K = np.eye(3, dtype=np.float32)
    p, n, c = ctx.cloud_from_depth(depth, None, K, None, False, 1, 0.1)
    assert len(p) == 0
    p, n, c = ctx.cloud_from_depth(depth, None, K, None, True, 1, 0.1)
    assert len(p) == 64 and np.isneginf(p).all() and np.isneginf(n).all()
    with pytest.raises(capi.IcpGpuError):
        ctx.cloud_from_depth(depth, None, K, np.zeros((4, 4), np.float32), False, 1, 0.1)      # singular extrinsics
    with pytest.raises(capi.IcpGpuError):
        ctx.cloud_from_depth(depth, None, K, None, False, 0, 0.1)                               # downsample 0


def test_projective_registration_from_depth_maps_equals_host_clouds(ctx):
    """reconstructRoom's path (main.cpp:183-341): both frames go up as depth maps and become target / source on the device."""
    w, h = 160, 120
    src, tgt, K, gt = synth.tum_pair(seed=9, width=w, height=h)
    room = synth.make_room(9)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    pos0 = np.array([3.0, 5.0, 1.4]); pos1 = pos0 + 10 * np.array([0.0015, 0.0006, 0.0002])
    d0, _ = synth.render_depth(room, pos0, 10.0, 0.0, w, h, fx, fy, cx, cy, seed=9, dropout=0.15)
    d1, _ = synth.render_depth(room, pos1, 11.0, 0.0, w, h, fx, fy, cx, cy, seed=10, dropout=0.15)
    c = capi.default_config()
    c.metric, c.matching, c.weighting, c.n_iterations, c.max_distance_sq = 2, 1, 2, 8, 0.1
    ctx.set_config(c); ctx.set_camera(K, w, h)
    ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
    pose_a, hist_a, n_a = ctx.estimate_pose()
    assert ctx.cloud_from_depth(d0, None, K, None, True, 1, 0.1, role=0, download=False) == w * h
    assert ctx.cloud_from_depth(d1, None, K, None, True, 1, 0.1, role=1, download=False) == w * h
    pose_b, hist_b, n_b = ctx.estimate_pose()
    assert n_a == n_b == 8 and np.array_equal(pose_a, pose_b) and np.array_equal(hist_a, hist_b)


@pytest.mark.parametrize("minimizer,metric", [(0, 1), (1, 2)])
def test_convergence_errors_bunny(ctx, bunny, minimizer, metric):
    src, tgt, gs, gt = bunny
    c = capi.default_config()
    c.metric, c.minimizer, c.n_iterations, c.max_distance_sq = metric, minimizer, 12, 0.0003
    ctx.set_config(c)
    ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
    ctx.set_correspondences(src.points[gs], tgt.points[gt])
    pose, hist, n_it = ctx.estimate_pose()
    rmse, bench = ctx.convergence_errors(benchmark=True)
    assert len(rmse) == len(bench) == n_it == 12
    for k in range(n_it):
        # the reference accumulates the 4 squared distances in fp32, the device in fp64: 1 ulp of the float result
        assert rmse[k] == pytest.approx(orc.rmse(hist[k], src.points[gs], tgt.points[gt]), rel=3e-7)
        assert bench[k] == pytest.approx(orc.benchmark_error(hist[k], src.points[gs], tgt.points[gt]), rel=1e-12)
    assert rmse[-1] < rmse[0]


def test_convergence_errors_all_points(ctx, small_eth_pair):
    """ETH / TUM runs use every point as a known correspondence (experiment.cpp): M = N, with non-finite points skipped."""
    src, tgt, gt_pose = small_eth_pair
    ref = orc.transform_points(gt_pose, src.points)
    s = src.points.copy(); s[17, 1] = np.nan; s[4000, 0] = -np.inf
    c = capi.default_config()
    c.metric, c.n_iterations, c.max_distance_sq = 1, 6, 0.5
    ctx.set_config(c)
    ctx.set_target(tgt.points, tgt.normals, tgt.colors); ctx.set_source(src.points, src.normals, src.colors)
    ctx.set_correspondences(s, ref)
    pose, hist, n_it = ctx.estimate_pose()
    rmse, _ = ctx.convergence_errors()
    assert len(rmse) == n_it == 6
    for k in range(n_it):
        assert rmse[k] == pytest.approx(orc.rmse(hist[k], s, ref), rel=2e-5)      # fp32 sequential sum of ~10^4 terms in the reference
    ok = np.isfinite(s).all(1)
    rmse2, bench2 = None, None
    ctx.set_correspondences(s[ok], ref[ok])
    rmse2, bench2 = ctx.convergence_errors(benchmark=True)
    for k in range(n_it):
        assert bench2[k] == pytest.approx(orc.benchmark_error(hist[k], s[ok], ref[ok]), rel=1e-9)


def test_convergence_errors_need_correspondences(bunny):
    src, tgt, _, _ = bunny
    with capi.Context(0) as c:
        c.set_target(tgt.points, tgt.normals, tgt.colors); c.set_source(src.points, src.normals, src.colors)
        c.estimate_pose()
        with pytest.raises(capi.IcpGpuError):
            c.convergence_errors()


def test_optimizer_mirror_fills_convergence_measure_from_the_device(bunny):
    """The reference-style driver flow (main.cpp:43-181) through the Python mirror of the operator API."""
    from icp_variants_b200.optimizer import ConvergenceMeasure, LinearICPOptimizer, TimeMeasure
    src, tgt, gs, gt = bunny
    opt = LinearICPOptimizer(device=0)
    opt.setMatchingMethod(0); opt.setMatchingMaxDistance(0.0003); opt.setMetric(2); opt.setNbOfIterations(20)
    cm = ConvergenceMeasure(src.points[gs], tgt.points[gt], runBenchmark=True); tm = TimeMeasure()
    opt.setConvergenceMeasure(cm); opt.setTimeMeasure(tm)
    pose = opt.estimatePose(src, tgt, np.eye(4, dtype=np.float32))
    assert len(cm.rmseErrors) == len(cm.benchmarkErrors) == 20
    assert cm.rmseErrors[-1] == pytest.approx(orc.rmse(pose, src.points[gs], tgt.points[gt]), rel=1e-6)
    assert 1.5e-4 < cm.rmseErrors[-1] < 2.5e-4 and tm.nIterations == 20


def test_optimizer_mirror_depth_frames(ctx):
    from icp_variants_b200.optimizer import LinearICPOptimizer
    w, h = 96, 72
    depth, rgbx, K = _frame(w, h, 5)
    opt = LinearICPOptimizer(device=0)
    assert opt.setTargetFromDepth(depth, rgbx, K, None, True) == w * h
    n = opt.setSourceFromDepth(depth, rgbx, K, None, False, 2)
    assert n == len(orc.cloud_from_depth(depth, rgbx, K[0, 0], K[1, 1], K[0, 2], K[1, 2], None, False, 2, 0.1)[0])


@pytest.mark.parametrize("projective,multires,minimizer", [(True, False, 0), (False, False, 0), (False, True, 1)])
def test_sequence_driver_equals_oracle_loop(projective, multires, minimizer):
    """reconstructRoom (main.cpp:183-341): target fixed at frame 0, pose carried over from frame to frame, RMSE against the
    ground-truth trajectory after every iteration -- the device pipeline against the same loop written with the oracle."""
    from icp_variants_b200.optimizer import CeresICPOptimizer, LinearICPOptimizer
    from icp_variants_b200.sequence import reconstructRoom
    w, h, n_it = 160, 120, 6
    frames, K, gt = synth.tum_sequence(n_frames=4, seed=21, width=w, height=h)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    opt = (CeresICPOptimizer if minimizer else LinearICPOptimizer)(device=0)
    opt.setMetric(2); opt.setNbOfIterations(n_it)
    if projective:
        opt.setMatchingMethod(1)
    opt.setMatchingMaxDistance(0.1)
    opt.setWeightingMethod(2 if projective else 0)
    opt.enableMultiResolution(multires)
    res = reconstructRoom(opt, frames, K, groundTruthPoses=gt)
    # the same loop with the oracle
    tp, tn, tc = orc.cloud_from_depth(frames[0], None, fx, fy, cx, cy, None, projective, 1, 0.1)
    cur = np.eye(4, dtype=np.float32)
    for i in range(1, 4):
        sp, sn, sc = orc.cloud_from_depth(frames[i], None, fx, fy, cx, cy, None, multires, 1 if multires else 8, 0.1)
        cfg = orc.Config(metric=2, minimizer=minimizer, matching=int(projective), weighting=2 if projective else 0, multires=multires,
                         n_iterations=n_it, max_distance_sq=0.1, fx=fx, fy=fy, cx=cx, cy=cy, width=w, height=h)
        rc, cur, hist, _ = orc.estimate_pose(cfg, sp, sn, sc, tp, tn, tc, init_pose=cur)
        assert rc == 0
        assert res.nSourcePoints[i - 1] == len(sp)
        got = res.cameraToWorld[i - 1]
        rot = 2.0 * np.arcsin(min(1.0, np.linalg.norm(got[:3, :3].astype(np.float64) - cur[:3, :3]) / (2.0 * np.sqrt(2.0))))
        assert rot < 1e-5 and np.abs(got[:3, 3] - cur[:3, 3]).max() < 1e-5, (i, rot)
        ref = orc.transform_points(gt[i], sp)
        r = np.array([orc.rmse(hh, sp, ref) for hh in hist])
        assert len(res.rmsePerIteration[i - 1]) == len(hist)
        assert np.allclose(res.rmsePerIteration[i - 1], r, rtol=3e-5)
        cur = got      # continue from the device's pose so that rounding differences do not accumulate across frames
    assert np.isfinite(res.finalRMSE).all() and len(res.estimatedPoses) == 4


def test_pair_queue_equals_sequential(small_eth_pair, bunny):
    """alignETH's pair loop (main.cpp:411-498) as a queue over two contexts of one GPU: same poses as one pair at a time,
    ConvergenceMeasure(source, unchanged source, runBenchmark=true) filled per pair (main.cpp:439)."""
    from icp_variants_b200.sequence import alignPairs
    src, tgt, pert = small_eth_pair
    bs, bt, _, _ = bunny
    unchanged = orc.transform_points(np.linalg.inv(pert).astype(np.float32), src.points)
    pairs = [(src, tgt, unchanged), (bs, bt), (src, tgt, unchanged), (tgt, src), (bs, bt)]
    cfg = capi.default_config()
    cfg.metric, cfg.n_iterations, cfg.max_distance_sq = 1, 8, 10.0
    with capi.Context(0) as a, capi.Context(0) as b:
        res = alignPairs([a, b], pairs, cfg, calculateErrors=True)
    assert len(res) == 5 and all(r is not None for r in res)
    with capi.Context(0) as c:
        for k, pr in enumerate(pairs):
            c.set_config(cfg); c.set_target(pr[1].points, pr[1].normals, pr[1].colors); c.set_source(pr[0].points, pr[0].normals, pr[0].colors)
            try:
                pose, hist, n_it = c.estimate_pose()
            except capi.IcpGpuError:
                assert res[k].error is not None
                continue
            assert np.array_equal(res[k].pose, pose) and res[k].nIterations == n_it
            if len(pr) > 2:
                assert len(res[k].rmseErrors) == len(res[k].benchmarkErrors) == n_it
                assert res[k].rmseErrors[-1] == pytest.approx(orc.rmse(pose, src.points, unchanged), rel=2e-5)
                assert res[k].benchmarkErrors[-1] == pytest.approx(orc.benchmark_error(pose, src.points, unchanged), rel=1e-9)


def test_pca_normals_bit_exact(ctx, small_eth_pair, bunny):
    """PointCloud(pcl cloud) (PointCloud.h:41-76): k = 5 PCA normals on the device against the oracle's restatement of
    pcl::NormalEstimation (itself equal, bit for bit, to the reference constructor running over the PCL stand-in)."""
    src, tgt, _ = small_eth_pair
    bs, bt, _, _ = bunny
    rng = np.random.default_rng(3)
    quant = (np.round(rng.uniform(-1, 1, (3000, 3)) * 16) / 16).astype(np.float32)       # many ties and duplicates
    holes = tgt.points.copy(); holes[11, 0] = np.nan; holes[500, 2] = -np.inf
    for pts, k, vp in ((tgt.points, 5, None), (bt.points, 5, None), (holes, 5, (1.0, -2.0, 0.5)), (quant, 5, None), (src.points, 8, None), (bt.points, 3, None)):
        ctx.set_target(pts, None, None)
        n_g, c_g = ctx.target_normals(k, vp, n=len(pts), curvature=True)
        n_o, c_o = orc.pca_normals(pts, k, vp if vp is not None else (0.0, 0.0, 0.0))
        assert np.array_equal(n_g, n_o, equal_nan=True), (k, np.nanmax(np.abs(n_g - n_o)), (n_g != n_o).any(1).sum())
        assert np.array_equal(c_g, c_o, equal_nan=True)
    assert np.isnan(n_g).sum() == 0


def test_pca_normals_become_the_target_normals(ctx, small_eth_pair):
    """The computed normals replace the target's: a point-to-plane registration then equals one with the normals uploaded."""
    src, tgt, _ = small_eth_pair
    c = capi.default_config(); c.metric, c.n_iterations, c.max_distance_sq = 1, 5, 0.5
    ctx.set_config(c)
    ctx.set_target(tgt.points, None, None); ctx.set_source(src.points, src.normals, None)
    nrm = ctx.target_normals(5, None, n=len(tgt.points))
    pose_a, _, _ = ctx.estimate_pose()
    ctx.set_target(tgt.points, nrm, None)
    pose_b, _, _ = ctx.estimate_pose()
    assert np.array_equal(pose_a, pose_b)


def test_pca_normals_full_size_properties(ctx):
    """370k points: unit length, oriented towards the viewpoint, and equal to an independent fp64 PCA (scipy k-NN + numpy
    eigh) on a random subset wherever the smallest eigenvalue is well separated."""
    from scipy.spatial import cKDTree
    _, tgt, _ = synth.eth_pair(seed=1234)
    pts = tgt.points
    ctx.set_target(pts, None, None)
    nrm = ctx.target_normals(5, None, n=len(pts))
    assert np.isfinite(nrm).all()
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-6)
    assert ((-pts * nrm).sum(1) >= -1e-6).all()                      # flipNormalTowardsViewpoint, viewpoint = origin
    sel = np.random.default_rng(0).choice(len(pts), 3000, replace=False)
    p64 = pts.astype(np.float64)
    _, nb = cKDTree(p64).query(p64[sel], k=5)
    ok = 0
    for s, idx in zip(sel, nb):
        q = p64[idx]; C = np.cov((q - q.mean(0)).T, bias=True)
        w, v = np.linalg.eigh(C)
        if w[1] - w[0] > 1e-3 * w[2]:
            assert abs(v[:, 0] @ nrm[s]) > 1 - 1e-6
            ok += 1
    assert ok > 1000
