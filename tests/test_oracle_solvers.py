"""Oracle stage 5/6 (metrics + minimisers) against literal restatements in numpy/scipy:
the explicit 4N x 6 systems of ICPOptimizer.h:676-898 solved by LAPACK SVD least squares (what
Eigen's JacobiSVD::solve computes), Kabsch via numpy SVD, and scipy's least_squares for the
Ceres functors of constraints.h."""
import numpy as np
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

from icp_variants_b200.synth import make_pose, rot_xyz
from oracle import oracle as O


def _matched_set(rng, m=400, noise=0.002):
    s = rng.uniform(-1, 1, size=(m, 3)).astype(np.float32)
    n = rng.normal(size=(m, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True)
    pose = make_pose([0.02, -0.01, 0.015], [1.0, -0.7, 1.5]).astype(np.float64)
    d = (s.astype(np.float64) @ pose[:3, :3].T + pose[:3, 3] + rng.normal(0, noise, size=(m, 3))).astype(np.float32)
    ns = (n @ np.linalg.inv(pose[:3, :3])).astype(np.float32)
    w = rng.uniform(0.2, 1.0, size=m).astype(np.float32)
    return s, d, ns.astype(np.float32), n.astype(np.float32), w, pose


def _literal_plane_system(s, d, n, w, dtype):
    s, d, n, w = (a.astype(dtype) for a in (s, d, n, w))
    m = len(s)
    A = np.zeros((4 * m, 6), dtype); b = np.zeros(4 * m, dtype)
    A[0::4, 0] = n[:, 2] * s[:, 1] - n[:, 1] * s[:, 2]
    A[0::4, 1] = n[:, 0] * s[:, 2] - n[:, 2] * s[:, 0]
    A[0::4, 2] = n[:, 1] * s[:, 0] - n[:, 0] * s[:, 1]
    A[0::4, 3:6] = n
    b[0::4] = (n * d).sum(1) - (n * s).sum(1)
    A[1::4, 1] = s[:, 2]; A[1::4, 2] = -s[:, 1]; A[1::4, 3] = 1; b[1::4] = d[:, 0] - s[:, 0]
    A[2::4, 0] = -s[:, 2]; A[2::4, 2] = s[:, 0]; A[2::4, 4] = 1; b[2::4] = d[:, 1] - s[:, 1]
    A[3::4, 0] = s[:, 1]; A[3::4, 1] = -s[:, 0]; A[3::4, 5] = 1; b[3::4] = d[:, 2] - s[:, 2]
    lam = np.empty(4 * m, dtype)
    lam[0::4] = dtype(1.0) * w
    for k in (1, 2, 3):
        lam[k::4] = dtype(np.float32(0.1)) * w
    return A * lam[:, None], b * lam


def test_point_to_plane_equals_literal_svd_least_squares():
    rng = np.random.default_rng(0)
    s, d, ns, n, w, _ = _matched_set(rng)
    rc, pose = O.solve_p2plane(s, d, n, w)
    assert rc == 0
    A, b = _literal_plane_system(s, d, n, w, np.float64)
    x = np.linalg.lstsq(A, b, rcond=None)[0]
    ref = np.eye(4); ref[:3, :3] = rot_xyz(*x[:3]); ref[:3, 3] = x[3:]
    assert np.allclose(pose, ref, atol=2e-7)
    # the reference's fp32 path (fp32 rows + fp32 SVD solve) stays within the north-star tolerance of it
    A32, b32 = _literal_plane_system(s, d, n, w, np.float32)
    x32 = np.linalg.lstsq(A32, b32, rcond=None)[0]
    assert np.abs(x32 - x).max() < 1e-5


def test_symmetric_equals_literal_system():
    rng = np.random.default_rng(1)
    s, d, ns, nt, w, _ = _matched_set(rng)
    rc, pose = O.solve_symmetric(s, d, ns, nt, w)
    assert rc == 0
    s64, d64, w64 = s.astype(np.float64), d.astype(np.float64), w.astype(np.float64)
    ms = np.float32(s64.mean(0)).astype(np.float64); md = np.float32(d64.mean(0)).astype(np.float64)
    sc, dc = s64 - ms, d64 - md
    nsum = nt.astype(np.float64) + ns.astype(np.float64)
    m = len(s)
    A = np.zeros((4 * m, 6)); b = np.zeros(4 * m)
    A[0::4, :3] = np.cross(sc + dc, nsum); A[0::4, 3:] = nsum; b[0::4] = ((dc - sc) * nsum).sum(1)
    A[1::4, 1] = sc[:, 2]; A[1::4, 2] = -sc[:, 1]; A[1::4, 3] = 1; b[1::4] = dc[:, 0] - sc[:, 0]
    A[2::4, 0] = -sc[:, 2]; A[2::4, 2] = sc[:, 0]; A[2::4, 4] = 1; b[2::4] = dc[:, 1] - sc[:, 1]
    A[3::4, 0] = sc[:, 1]; A[3::4, 1] = -sc[:, 0]; A[3::4, 5] = 1; b[3::4] = dc[:, 2] - sc[:, 2]
    lam = np.empty(4 * m); lam[0::4] = w64
    for k in (1, 2, 3):
        lam[k::4] = np.float64(np.float32(0.1)) * w64
    A *= lam[:, None]; b *= lam
    M = A.T @ A + np.float64(np.float32(0.0001) * np.float32(0.0001)) * np.eye(6)
    x = np.linalg.solve(M, A.T @ b)
    a_t, t_t = x[:3], x[3:]
    tan = np.linalg.norm(a_t); a = a_t / tan
    sin = tan / np.sqrt(1 + tan * tan); cos = sin / tan
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + sin * K + (1 - cos) * K @ K
    def T(v):
        t = np.eye(4); t[:3, 3] = v; return t
    R4 = np.eye(4); R4[:3, :3] = R
    ref = T(md) @ R4 @ T(t_t * cos) @ R4 @ T(-ms)
    assert np.allclose(pose, ref, atol=3e-6)


def test_point_to_point_is_weighted_kabsch():
    rng = np.random.default_rng(2)
    s, d, _, _, w, true_pose = _matched_set(rng, noise=0.0)
    rc, pose = O.solve_p2p(s, d, w)
    assert rc == 0
    s64, d64 = s.astype(np.float64), d.astype(np.float64)
    sm, dm = s64.mean(0), d64.mean(0)
    H = (d64 - dm).T @ (w.astype(np.float64)[:, None] * (s64 - sm))   # ProcrustesAligner.h:50-54
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.linalg.det(U @ Vt)])
    R = U @ D @ Vt
    t = R @ (dm - sm) - R @ dm + dm
    assert np.allclose(pose[:3, :3], R, atol=1e-6) and np.allclose(pose[:3, 3], t, atol=1e-6)
    assert np.allclose(pose, true_pose, atol=1e-5)      # noise-free => exact recovery


def test_point_to_point_reflection_case():
    # planar, mirrored configuration: det(U V^T) < 0 must be fixed on the smallest singular direction
    rng = np.random.default_rng(3)
    s = rng.uniform(-1, 1, size=(50, 3)).astype(np.float32); s[:, 2] = 0
    d = s.copy(); d[:, 0] *= -1
    rc, pose = O.solve_p2p(s, d, np.ones(50, np.float32))
    assert rc == 0 and np.isclose(np.linalg.det(pose[:3, :3].astype(np.float64)), 1.0, atol=1e-5)


def test_no_matches_is_an_error_not_a_hang():
    z = np.zeros((0, 3), np.float32)
    assert O.solve_p2p(z, z, np.zeros(0, np.float32))[0] == -1
    assert O.solve_p2plane(z, z, z, np.zeros(0, np.float32))[0] == -1
    assert O.solve_symmetric(z, z, z, z, np.zeros(0, np.float32))[0] == -1


def _ceres_residuals(x, metric, sp, sn, d, n, w):
    R = Rotation.from_rotvec(x[:3]).as_matrix()
    y = sp @ R.T + x[3:]
    r = [np.float64(np.float32(0.1)) * w[:, None] * (y - d)]
    if metric == 1:
        r.append((w * ((y - d) * n).sum(1))[:, None])
    if metric == 2:
        z = d @ R                                       # R(-omega) d = R^T d
        r.append((w * ((y - z) * (n + sn)).sum(1))[:, None])
    return np.concatenate(r, 1).ravel()


def test_lm_restatement_reaches_the_least_squares_optimum():
    rng = np.random.default_rng(4)
    s, d, ns, nt, w, _ = _matched_set(rng, m=300)
    matches = np.zeros(len(s), O.MATCH_DTYPE); matches["idx"] = np.arange(len(s)); matches["weight"] = w
    for metric in (0, 1, 2):
        rc, x, pose, nlm = O.solve_lm(metric, s, ns, d, nt, matches, max_iterations=50)
        assert rc == 0 and 1 <= nlm <= 50
        args = (metric, s.astype(np.float64), ns.astype(np.float64), d.astype(np.float64), nt.astype(np.float64), w.astype(np.float64))
        sol = least_squares(_ceres_residuals, np.zeros(6), args=args, method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14)
        # Ceres stops on |dcost| <= 1e-6*cost WITHOUT applying the candidate step, so the result sits
        # within ~1e-6..1e-5 of the true optimum rather than at it.
        assert np.allclose(x, sol.x, atol=1e-5), (metric, x, sol.x)
        # with Ceres' max_num_iterations = 10 the result is already within tolerance on this well-posed case
        rc, x10, _, n10 = O.solve_lm(metric, s, ns, d, nt, matches, max_iterations=10)
        assert n10 <= 10 and np.allclose(x10, sol.x, atol=1e-5)
        # PoseIncrement::convertToMatrix
        assert np.allclose(pose[:3, :3], Rotation.from_rotvec(x[:3]).as_matrix(), atol=1e-6)
        assert np.allclose(pose[:3, 3], x[3:], atol=1e-7)


def test_lm_skips_plane_row_for_non_finite_normal():
    rng = np.random.default_rng(5)
    s, d, ns, nt, w, _ = _matched_set(rng, m=100)
    nt2 = nt.copy(); nt2[::3] = -np.inf
    matches = np.zeros(len(s), O.MATCH_DTYPE); matches["idx"] = np.arange(len(s)); matches["weight"] = w
    rc, x, pose, _ = O.solve_lm(1, s, ns, d, nt2, matches)
    assert rc == 0 and np.isfinite(pose).all()
    rc, pose2 = O.solve_p2plane(s, d, nt2, w)
    assert rc == 0 and np.isfinite(pose2).all()      # documented deviation from the linear reference (NaN pose)
